"""EnKF state bridge: state vector of extract_sub_ and the analysis update of enkf_wghmstate_.
CPU: the numpy restatement against the COMPILED REFERENCE's extract_sub_ / enkf_wghmstate_ / setStorages (golden fixture
tests/golden/ref_ng1000_enkf.npz, bit for bit) and on hand-checked cells.  GPU: the CUDA kernels against the restatement
bit for bit, and against the reference golden through a 31-day run."""
import os

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
STORAGES = ["canopy", "snow", "soil", "loc_lake_stor", "loc_wetl_stor", "glo_lake_stor", "glo_wetl_stor", "res_stor", "river_stor", "gw"]


@pytest.fixture(scope="module")
def golden_enkf():
    z = np.load(os.path.join(ROOT, "tests", "golden", "ref_ng1000_enkf.npz"))
    return {k: z[k] for k in z.files}


def test_restatement_bit_exact_vs_reference_golden(golden, golden_enkf):
    """oracle/enkf_bridge.py against what the reference's own functions computed for January 1901 of the 1000-cell
    world: daily routing entries (routing.cpp:5002-5020), Cell::mean, extract_sub_, enkf_wghmstate_ (incl. the snow
    in elevation rescale) and the restore by setStorages - every value bit-identical."""
    from oracle import enkf_bridge as eb
    g = golden_enkf
    cells = g["cells"]
    n = cells.size
    st = {k[7:]: g[k] for k in g if k.startswith("before/") and k != "before/snow_bands"}
    st["area"], st["contfreq"] = golden["d0/area"][cells], golden["d0/contfreq"][cells]
    last = eb.daily_entry(st)
    assert np.array_equal(last[:, 3:], g["routing_mm"][30].T)
    days = []
    for d in range(31):
        e = np.zeros((n, 10))
        e[:, 3:] = g["routing_mm"][d].T
        days.append(e)
    mon = eb.monthly_mean(days, last)
    assert np.array_equal(mon, g["enkf_month_mean"])
    idx = np.arange(n)
    assert np.array_equal(eb.extract_sub(mon, idx, g["mean_field"]), g["enkf_extract"])
    st1, sb1 = eb.enkf_update(st, g["before/snow_bands"], idx, mon, g["enkf_field"], g["enkf_prediction"], g["mean_field"])
    for k in STORAGES:
        assert np.array_equal(st1[k], g["after/" + k]), k
    assert np.array_equal(sb1, g["after/snow_bands"])
    # the fixture reaches the limits of enKF2wghmState.cpp:89-121 / 440-471
    inc = g["enkf_field"] - g["enkf_prediction"]
    assert (g["enkf_lastday"][:, 1] == 1000.).sum() >= 3
    assert ((g["enkf_lastday"][:, [0, 2, 4, 6, 7, 8]] == 0.) & (inc[:, [0, 2, 4, 6, 7, 8]] < -40)).sum() >= 10  # emptied by the analysis
    assert (g["enkf_lastday"][:, [3, 5, 9]] < 0.).any()                      # lakes and groundwater may go negative
    assert (g["enkf_month_mean"][:, 1] == 0.).any() and (g["enkf_month_mean"][:, 1] > 0.).sum() > 20  # both snow branches
    assert (g["enkf_snow_elev"][:, 1:] == 1000.).any()

FIELDS = ["canopy", "snow", "soil", "loc_lake_stor", "loc_wetl_stor", "glo_lake_stor", "glo_wetl_stor", "res_stor", "river_stor", "gw",
          "land_area_frac", "land_area_frac_next", "status_laf_next"]


def test_restatement_on_two_cells():
    from oracle import enkf_bridge as eb
    st = {"area": np.array([3000., 2500.]), "contfreq": np.array([100., 50.]), "status_laf_next": np.array([1, 1]),
          "land_area_frac": np.array([90., 40.]), "land_area_frac_next": np.array([80., 0.]),
          "canopy": np.array([1., 2.]), "snow": np.array([10., 5.]), "soil": np.array([100., 50.])}
    for k in eb.ROUTING:
        st[k] = np.array([0.03, 0.0125])
    v = eb.daily_entry(st)
    assert v[0, 0] == 1. * 80. / 100. and v[0, 3] == 0.03 / ((3000. * 1.) / 1e6) and v[1, 1] == 0.
    mon = eb.monthly_mean([v, v, v], v)
    assert np.allclose(mon, v, rtol=1e-15)
    sb = np.zeros((2, 101))
    sb[:, 1:] = [[10.], [5.]]
    zero = np.zeros((2, 10))
    # analysis == prediction and unchanged monthly snow: nothing may move except the band rescale factor of exactly 1
    f = eb.extract_sub(mon, [0, 1])
    st2, sb2 = eb.enkf_update(st, sb, [0, 1], mon, f, f, zero)
    assert st2["canopy"][0] == v[0, 0] * 100. / 80. and st2["canopy"][1] == 0. and np.array_equal(sb2[0], sb[0]) and (sb2[1] == 0).all()
    # limits: a strongly negative increment empties the bounded compartments, lakes and groundwater go negative
    st3, _ = eb.enkf_update(st, sb, [0], mon, f[:1] - 1e6, f[:1], zero[:1])
    assert st3["canopy"][0] == 0. and st3["river_stor"][0] == 0. and st3["gw"][0] < 0. and st3["loc_lake_stor"][0] < 0.
    st4, sb4 = eb.enkf_update(st, sb, [0], mon, f[:1] + 1e6, f[:1], zero[:1])
    assert st4["snow"][0] == 1000. * 100. / 80. and sb4[0, 1:].max() == 1000. * 100. / 80.


@pytest.mark.gpu
def test_gpu_bridge_vs_reference_golden(golden, golden_enkf):
    """the CUDA path on the reference's month: 31 days from the golden cold start with wgk_month_begin, then
    wgk_state_vector against extract_sub_'s vector and wgk_enkf_update against the state the reference holds after
    enkf_wghmstate_ + setStorages (free-run tolerance policy of DESIGN.md 6; the bit-exact link is restatement <-> reference
    above and CUDA <-> restatement below)"""
    import watergap2_b200 as wg
    from tests.util import assert_parity, golden_day
    g = golden_enkf
    ng = int(golden["ng"])
    d0 = golden_day(golden, 0)
    ro = np.zeros(ng, np.int32)
    ro[d0["routing_cell"] - 1] = np.arange(1, ng + 1)
    m = wg.Model(ng)
    m.set_topology(ro, d0["downstream_cell"])
    m.load(d0)
    m.forcing_reserve(31)
    f = {k: golden[f"forcing1/{k}"] for k in ("P", "T", "SW", "LW")}
    m.set_forcing(0, 31, f["P"], f["T"], f["SW"], f["LW"])
    m.month_begin()
    m.step_days(1, 0, 1, 0, 31)
    cells = g["cells"]
    flips = max(1, cells.size // 50)
    got = m.state_vector(cells, "month", mean_field=g["mean_field"])
    for k, name in enumerate(STORAGES):  # mm over the continental area
        assert_parity("snow", g["enkf_extract"][:, k] + g["mean_field"][:, k], got[:, k] + g["mean_field"][:, k], max_flips=flips)
    m.enkf_update(cells, g["enkf_field"], g["enkf_prediction"], g["mean_field"])
    for name in STORAGES:
        assert_parity(name, g["after/" + name], m.get(name)[cells], max_flips=flips)
    sb = m.get("snow_bands").reshape(-1, 101)[cells]
    assert_parity("snow_bands", g["after/snow_bands"][:, 1:].ravel(), sb[:, 1:].ravel(), max_flips=flips * 100)


@pytest.mark.gpu
def test_gpu_bridge_matches_restatement(world3000):
    from oracle import enkf_bridge as eb, synth_world as sw, wg_init
    import watergap2_b200 as wg
    w = world3000
    ini = wg_init.derive(w)
    topo = ini["_topology"]
    f = sw.forcing_month(w, 1901, 1)

    def model():
        m = wg.Model(w.ng, nmember=2)
        m.set_topology(topo["rout_order"], topo["outflow_cell"], cell_class=wg.cell_classes(ini))
        m.load(ini)
        m.forcing_reserve(31)
        m.set_forcing(0, 31, f["P"], f["T"], f["SW"], f["LW"])
        return m

    def state(m, mem):
        st = {k: m.get(k, mem) for k in FIELDS}
        st["area"], st["contfreq"] = np.asarray(ini["area"], np.float64), np.asarray(ini["contfreq"], np.float64)
        return st

    m = model()
    m.step_days(1, 0, 1, 0, 20)  # spin up some storage and snow before the "month" starts
    m.month_begin()
    days = []
    for d in range(21, 27):  # six single-day calls, then three days in one graph
        m.step_days(d, 0, d, d - 1, 1)
        days.append(eb.daily_entry(state(m, 1)))
    m2 = model()  # the same 29 days with the last three of the month in one call
    m2.step_days(1, 0, 1, 0, 20)
    m2.month_begin()
    m2.step_days(21, 0, 21, 20, 3)
    m2.step_days(24, 0, 24, 23, 3)
    cells = np.arange(5, w.ng, 7, dtype=np.int32)
    mon = eb.monthly_mean(days, days[-1])
    rng = np.random.default_rng(3)
    mean_field = rng.normal(0., 5., (cells.size, 10))
    got = m.state_vector(cells, "month", member=1, mean_field=mean_field)
    assert np.array_equal(got, eb.extract_sub(mon, cells, mean_field))
    assert np.array_equal(m2.state_vector(cells, "month", member=1, mean_field=mean_field), got)
    assert np.array_equal(m.state_vector(cells, "lastday", member=1), days[-1][cells])
    assert (got[:, 1] + mean_field[:, 1] > 0).sum() > 20  # snow present in the region

    # analysis step: random increments, some of them large enough to hit every limit
    pred = got
    field = pred + rng.normal(0., 3., pred.shape) * (rng.random(pred.shape) < 0.8) - 50. * (rng.random(pred.shape) < 0.05)
    st0, sb0 = state(m, 1), m.get("snow_bands", 1)
    other = {k: m.get(k, 0) for k in FIELDS + ["snow_bands"]}
    m.enkf_update(cells, field, pred, mean_field, member=1)
    st1, sb1 = eb.enkf_update(st0, sb0, cells, mon, field, pred, mean_field)
    for k in FIELDS:
        assert np.array_equal(m.get(k, 1), st1[k]), k
    assert np.array_equal(m.get("snow_bands", 1).reshape(-1, 101)[:, 1:], sb1[:, 1:])
    for k, v in other.items():  # the other member is untouched
        assert np.array_equal(m.get(k, 0), v), k
    # the updated member keeps running (the snow-free bookkeeping is rebuilt) and equals a fresh context
    # started from the restated update
    m.step_days(27, 0, 27, 26, 2)
    m3 = model()
    for k in FIELDS:
        m3.set(k, st1[k], 1)
        m3.set(k, m2.get(k, 0), 0)
    m3.set("snow_bands", sb1, 1)
    for k in ("lai_days", "lai_status", "lai_precsum", "red_loc_lake", "red_loc_wetl", "red_glo_lake", "red_glo_wetl", "red_res", "red_river",
              "k_release", "land_area_frac_prev", "fswb_laf", "fswb_laf_next", "river_area_frac_next", "river_area_frac_change",
              "storage_transfer"):
        m3.set(k, m2.get(k, 1), 1)
    m3.step_days(27, 0, 27, 26, 2)
    for k in FIELDS + ["discharge"]:
        assert np.array_equal(m.get(k, 1), m3.get(k, 1)), k
