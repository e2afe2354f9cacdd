"""TEST INFRASTRUCTURE - numpy restatement of the state exchange of the reference's PDAF coupling.

  daily_entry()    the ten WghmStateFile compartments of one day in mm over the continental area:
                   routing.cpp:5002-5020 (seven routing compartments, storage / ((area * (contfreq / 100.)) / 1000000.))
                   and integrateWGHM.cpp:838-847 (canopy, snow, soil * landAreaFrac / contfreq, the month-end value
                   written to every day of the month)
  monthly_mean()   Cell::mean (wghmStateFile.cpp:711-728): std::accumulate from 0.0 in day order, / number of days
  extract_sub()    extractsub.cpp:65-79: [ncells][10] minus the temporal mean field
  enkf_update()    enKF2wghmState.cpp:89-121 (last day += field - prediction, limits), :440-471 (snow bands) and the
                   restore of the next cycle (daily.cpp:1896-1924, routing.cpp:851-882)

Pinned against the compiled reference: `ref_harness replay --enkf` (oracle/ref_harness.cpp) calls the reference's own
extract_sub_, enkf_wghmstate_ and setStorages on January 1901 of the 1000-cell world; every function here is bit-identical
to what they computed (tests/golden/ref_ng1000_enkf.npz, made by tests/golden/make_golden_enkf.py; checked by
tests/test_enkf_bridge.py::test_restatement_bit_exact_vs_reference_golden).  Note enKF2wghmState.cpp:92-120 reads
`field[count] - prediction[count++]` (unsequenced); g++ 13 -O2 evaluates both at the same index, which is what is restated.
Only tests/ may import this module.
"""
import numpy as np

ROUTING = ["loc_lake_stor", "loc_wetl_stor", "glo_lake_stor", "glo_wetl_stor", "res_stor", "river_stor", "gw"]


def _laf(st):
    return np.where(st["status_laf_next"] == 0, st["land_area_frac"], st["land_area_frac_next"])


def _denom(st):
    return (st["area"] * (st["contfreq"] / 100.)) / 1000000.


def daily_entry(st):
    """st: dict of per-cell arrays (state + area, contfreq) -> [ncell, 10]"""
    laf, contf = _laf(st), st["contfreq"]
    v = np.empty((st["area"].size, 10))
    for k, name in enumerate(("canopy", "snow", "soil")):
        v[:, k] = st[name] * laf / contf
    d = _denom(st)
    for k, name in enumerate(ROUTING):
        v[:, 3 + k] = st[name] / d
    return v


def monthly_mean(days, month_end):
    """days: list of daily_entry() of every day of the month; canopy/snow/soil of every day are overwritten with the
    month-end values (integrateWGHM.cpp:843-847) before the mean is taken"""
    n = len(days)
    acc = np.zeros_like(month_end)
    for v in days:
        e = v.copy()
        e[:, :3] = month_end[:, :3]
        acc = acc + e
    return acc / float(n)


def extract_sub(vec, cells, mean_field=None):
    out = vec[cells].copy()
    return out if mean_field is None else out - mean_field


def enkf_update(st, snow_bands, cells, mon_mean, field, prediction, mean_field):
    """-> (updated state dict, updated snow_bands [ncell, 101]) for the cells of the region"""
    st = {k: np.array(v, copy=True) for k, v in st.items()}
    sb = np.array(snow_bands, copy=True).reshape(-1, 101)
    laf, contf, den = _laf(st), st["contfreq"], _denom(st)
    last = daily_entry(st)
    for j, n in enumerate(cells):
        w = last[n] + (field[j] - prediction[j])
        for k in (0, 1, 2, 4, 6, 7, 8):
            if w[k] < 0.:
                w[k] = 0.
        if w[1] > 1000.:
            w[1] = 1000.
        before, after = mon_mean[n, 1], field[j, 1] + mean_field[j, 1]
        for e in range(1, 101):
            sie = 0. if laf[n] == 0. else sb[n, e] * laf[n] / contf[n]
            sie = after / 100 if before == 0 else sie * (after / before)
            sie = min(max(sie, 0.), 1000.)
            sb[n, e] = 0. if laf[n] <= 0. else sie * contf[n] / laf[n]
        for k, name in enumerate(("canopy", "snow", "soil")):
            st[name][n] = 0. if laf[n] <= 0. else w[k] * contf[n] / laf[n]
        for k, name in enumerate(ROUTING):
            st[name][n] = w[3 + k] * den[n]
    return st, sb
