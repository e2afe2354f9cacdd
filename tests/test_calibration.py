"""Sweep objectives: the reference's criteria (calibration.cpp:300-330, 531-546) for all parameter sets."""
import numpy as np
import pytest

from watergap2_b200 import calibration as cal


def test_criteria_formulas():
    m = np.array([10., 12., -99., 8.], np.float32)
    s = np.array([11., 11., 50., 9.], np.float32)
    c = cal.criteria(m, s)
    avg = np.float32(30.) / np.float32(3.)
    assert c["years"] == 3 and c["measured_avg"] == float(avg) and c["sum_of_differences"] == 1.0
    sum1 = sum(np.float32(x - avg) ** 2 for x in (10., 12., 8.))
    assert abs(c["nse"] - float((sum1 - np.float32(3.)) / sum1)) < 1e-7
    assert cal.criteria(m[:1], s[:1])["nse"] == -99.0
    assert cal.criteria(m, m)["nse"] == 1.0 and cal.criteria(m, m)["rel_difference"] == 0.0
    q = cal.measured_km3_per_year([1000., -99.])
    assert q[0] == np.float32(float(np.float32(1000.) * np.float32(365) * np.float32(24) * np.float32(60) * np.float32(60)) / 1e9) and q[1] == -99
    rec = np.ones((730, 2)) * [[1e-3, 2e-3]]
    assert np.allclose(cal.annual_runoff_km3(rec), [[0.365, 0.73], [0.365, 0.73]], rtol=1e-6)
    assert abs(cal.correction_factor(10., 24., 0., 0., 2) - 10. / 12.) < 1e-6


@pytest.mark.gpu
def test_sweep_finds_the_generating_parameter_set(world3000):
    """four parameter sets differing in the runoff coefficient gamma as members of one context; the 'observations'
    are the annual discharge of set 2: its criteria must be perfect and the sweep must pick it"""
    from oracle import synth_world as sw, wg_init
    import watergap2_b200 as wg
    w = world3000
    psets = []
    for g in (0.5, 1.0, 2.0, 4.0):
        p = np.array(sw.default_params(w, 0), np.float64).reshape(26, -1).copy()
        p[0] = g
        psets.append(p)
    inits = [wg_init.derive(w, p) for p in psets]
    topo = inits[0]["_topology"]
    m = wg.Model(w.ng, nmember=4, npset=4)
    m.set_topology(topo["rout_order"], topo["outflow_cell"], cell_class=wg.cell_classes(inits[0]))
    for i, ini in enumerate(inits):
        m.load(ini, pset=i)
        m.set_member_pset(i, i)
    f = sw.forcing_month(w, 1901, 1)
    m.forcing_reserve(31)
    m.set_forcing(0, 31, f["P"], f["T"], f["SW"], f["LW"])
    stations = np.argsort(-w.acc)[:3].astype(np.int32)
    m.record_cells(stations, 62)
    m.step_days(1, 0, 1, 0, 62)  # two 31-day "years" on the January forcing (slots cycle)
    obs = cal.annual_runoff_km3(m.get_record(62, 2), days_per_year=31)
    crit = cal.sweep_criteria(m, obs, 2, days_per_year=31)
    assert crit[2][0]["nse"] == 1.0 and crit[2][0]["rel_difference"] == 0.0
    assert cal.best_member(crit, 0) == 2
    sims = [cal.annual_runoff_km3(m.get_record(62, k), 31)[:, 0].sum() for k in range(4)]
    assert sims[0] > sims[1] > sims[2] > sims[3]  # a larger gamma keeps more water in the soil
