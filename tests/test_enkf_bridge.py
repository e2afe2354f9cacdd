"""EnKF state bridge: state vector of extract_sub_ and the analysis update of enkf_wghmstate_.
CPU: the numpy restatement on hand-checked cells.  GPU: the CUDA kernels against it, bit for bit."""
import numpy as np
import pytest

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
