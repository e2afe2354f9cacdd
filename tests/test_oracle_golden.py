"""CPU: the oracle restatement against golden vectors dumped from the compiled reference."""
import os

import numpy as np
import pytest

from tests.util import golden_day


def _oracle_from_golden(wgo, golden):
    ng = int(golden["ng"])
    o = wgo.Oracle(ng)
    n = o.load_records({(k, 0): v for k, v in golden_day(golden, 0).items()}, 0)
    assert n > 70
    return o


def test_oracle_bit_exact_vs_reference_golden(golden, oracle_lib):
    """59 simulated days on the 1000-cell world: every state and flux array the reference holds
    in memory (incl. the 100 snow bands and per-cell river discharge) must be BIT-IDENTICAL."""
    wgo = oracle_lib
    o = _oracle_from_golden(wgo, golden)
    days = [int(d) for d in golden["days"]]
    curm = -1
    checked = 0
    for sd in range(1, max(days) + 1):
        doy, mon, dom = wgo.calendar(sd)
        if mon != curm:
            o.set_forcing_month({k: golden[f"forcing{mon + 1}/{k}"] for k in ("P", "T", "SW", "LW")})
            curm = mon
        o.step_day(doy, mon, dom)
        if sd in days:
            for name, ref in golden_day(golden, sd).items():
                if not o.has(name):
                    continue
                got = o.field(name)
                assert np.array_equal(ref, got), f"day {sd} field {name}: {int((ref != got).sum())} cells differ"
                checked += 1
    assert checked > 200
    # the fixture must exercise the interesting branches
    g59 = golden_day(golden, 59)
    assert (g59["discharge"] > 0).sum() > 500 and (g59["snow"] > 0).sum() > 100
    assert (g59["res_stor"] > 0).sum() >= 5 and (g59["glo_lake_stor"] != 0).sum() >= 5


def test_oracle_bit_exact_deep_snow(golden_deep, oracle_lib):
    """snow packs of up to 1400 mm per band: pins the 1000 mm cap (daily.cpp:958-976), sublimation
    and melt of deep packs; bands, sums and all downstream fluxes bit-identical for 6 days."""
    wgo = oracle_lib
    g = golden_deep
    o = _oracle_from_golden(wgo, g)
    assert (golden_day(g, 0)["snow_bands"] > 1000).sum() > 5000
    o.set_forcing_month({k: g[f"forcing1/{k}"] for k in ("P", "T", "SW", "LW")})
    checked = 0
    for sd in range(1, 7):
        o.step_day(sd, 0, sd)
        if sd in (1, 2, 6):
            for name, ref in golden_day(g, sd).items():
                if o.has(name):
                    assert np.array_equal(ref, o.field(name)), f"day {sd} field {name}: {int((ref != o.field(name)).sum())} differ"
                    checked += 1
    assert checked > 100


def test_topology_bit_exact_vs_reference_files(golden, oracle_lib, world1000):
    """rout_prepare.cpp restatement against the routing files the reference wrote."""
    wgo = oracle_lib
    w = world1000
    t = wgo.rout_prepare(w.flowdir, w.row, w.col, w.gcrc.T)
    for fname, key in [("G_LDD_2.UNF1", "ldd_2"), ("G_INFLC.9.UNF4", "inflow9"), ("G_FLOW_ACC.UNF2", "flow_acc"),
                       ("G_CELLS_TO_OUTLET.UNF2", "cells_to_outlet"), ("G_BASINS.UNF2", "basins"),
                       ("G_BASINS_2.UNF2", "basins2"), ("G_OUTFLC.UNF4", "outflow_cell"),
                       ("G_ROUT_ORDER.UNF4", "rout_order")]:
        assert np.array_equal(golden["routing/" + fname], t[key].ravel()), fname
    cd, slope, length = wgo.river_geometry(w.altitude, w.meander, t["outflow_cell"], t["ldd"], w.row, w.col)
    assert np.array_equal(golden["routing/G_RIVERSLOPE.UNF0"], slope)
    assert np.array_equal(golden["routing/G_RIVER_LENGTH.UNF0"], length)
    alloc, sm = wgo.reservoir_prepare(w.resarea, w.mean_outflow, w.mean_outflow12, t["outflow_cell"])
    assert np.array_equal(golden["routing/G_ALLOC_COEFF.5.UNF0"], alloc.ravel())
    assert np.array_equal(golden["routing/G_START_MONTH.UNF1"], sm)
    # the routing order is a level-major topological order
    ro, down = t["rout_order"], t["outflow_cell"]
    has = down > 0
    assert (ro[has] < ro[down[has] - 1]).all()
    assert sorted(ro.tolist()) == list(range(1, w.ng + 1))


def test_init_restatement_bit_exact(golden, oracle_lib, world1000):
    """gw_frac / s_max / lai.init / routing.init / annualInit / setLakeWetlToMaximum restated
    in oracle/wg_init.py against the reference's in-memory values before the first day."""
    from oracle import wg_init
    d = wg_init.derive(world1000)
    n = 0
    for name, ref in golden_day(golden, 0).items():
        if name not in d:
            continue
        got = np.asarray(d[name]).ravel().astype(ref.dtype)
        assert np.array_equal(ref, got), f"{name}: {int((ref != got).sum())} cells differ"
        n += 1
    assert n >= 75


def test_empty_and_edge_inputs(oracle_lib):
    """single-cell world and a two-cell chain: no upstream, sink handling, zero forcing."""
    wgo = oracle_lib
    o = wgo.Oracle(2)
    for k in ("area", "contfreq"):
        o.set(k, [3000.0, 3000.0] if k == "area" else [100.0, 100.0])
    o.set("contcell", [1, 1]); o.set("toBeCalculated", [1, 1]); o.set("landcover", [9, 9])
    o.set("ldd", [6, 5]); o.set("downstream_cell", [2, 0]); o.set("routing_cell", [1, 2])
    o.set("land_area_frac", [100.0, 100.0]); o.set("river_length", [50.0, 55.0]); o.set("river_slope", [1e-3, 1e-3])
    o.set("roughness", [0.05, 0.05]); o.set("river_bottom_width", [10.0, 10.0]); o.set("river_width_bf", [20.0, 20.0])
    o.set("river_storage_max", [1e-3, 1e-3]); o.set("smax", [100.0, 100.0]); o.set("gwfactor", [0.5, 0.5])
    o.set("rgmax", [700, 700]); o.set("texture", [20, 20]); o.set("laimax", [2.0, 2.0])
    p = np.zeros((26, 2)); 
    from oracle import synth_world as sw
    p[:] = np.array(sw.PARAM_DEFAULT)[:, None]
    o.set("params", p); o.set("gamma_hbv", p[0]); o.set("cfa", p[1]); o.set("cfs", p[2])
    o.set("lai_kc_min", np.full(18, 0.4)); o.set("lai_kc_max", np.full(18, 1.0)); o.set("lct_emissivity", np.full(18, 0.98))
    o.set("lct_ddf", np.full(18, 3.0)); o.set("lct_albedo_snow", np.full(18, 0.4)); o.set("lai_initial_days", np.full(18, 10))
    o.set("lai_factor_a", np.full(18, 0.05)); o.set("lai_factor_b", np.full(18, 0.4))
    z = np.zeros((2, 31), np.float32)
    o.set_forcing_month({"P": z + 5, "T": z + 15, "SW": z + 200, "LW": z + 350})
    for d in range(1, 11):
        o.step_day(d, 0, d)
    q = o.field("discharge")
    assert q[0] > 0 and q[1] > q[0] * 0.5 and np.isfinite(o.total_storage_km3())
    # dry, frozen world: nothing moves, nothing goes negative or NaN
    o.set_forcing_month({"P": z, "T": z - 30, "SW": z, "LW": z + 150})
    for d in range(11, 21):
        o.step_day(d, 0, d)
    for k in ("soil", "canopy", "snow", "river_stor", "gw"):
        v = o.field(k)
        assert np.isfinite(v).all() and (v >= 0).all(), k


def test_wateruse_fixture_is_a_different_run(golden):
    """groundwork for SURVEY 8f-4 (water use; not on the product path yet): the fixture made by
    tests/golden/make_golden_wateruse.py holds the compiled reference's run with net abstractions - same world and cold start as
    ref_ng1000.npz, different storages from the first day on, use left unsatisfied, groundwater depleted below zero"""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_ng1000_wateruse.npz"))
    assert int(z["ng"]) == int(golden["ng"]) and {1, 2, 31, 59} <= set(int(d) for d in z["days"])
    assert np.array_equal(z["d1/canopy"], golden["d1/canopy"])              # the vertical balance does not see the water use
    assert not np.array_equal(z["d59/river_stor"], golden["d59/river_stor"])
    assert (z["d59/wu_total_unsatisfied"] > 0).sum() > 50 and (z["d59/gw"] < 0).any()
    assert z["input/G_NETUSE_SW_m3_1901.12.UNF0"].size == 12 * int(z["ng"])


def test_wateruse_daily_abstraction_bit_exact(golden):
    """oracle/water_use.py::daily_net_abstraction against G_dailydailyNUs / G_dailydailyNUg of the compiled reference (January and
    February; the groundwater value of the first day, before updateNetAbstractionGW adapts it)"""
    from oracle import water_use as wu
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_ng1000_wateruse.npz"))
    par = golden["d0/params"].reshape(26, -1)
    sw_m3, gw_m3 = z["input/G_NETUSE_SW_m3_1901.12.UNF0"], z["input/G_NETUSE_GW_m3_1901.12.UNF0"]
    assert np.array_equal(wu.daily_net_abstraction(sw_m3, par[23], 0), z["d1/wu_daily_nus"])
    assert np.array_equal(wu.daily_net_abstraction(sw_m3, par[23], 0), z["d31/wu_daily_nus"])
    assert np.array_equal(wu.daily_net_abstraction(sw_m3, par[23], 1), z["d59/wu_daily_nus"])
    assert np.array_equal(wu.daily_net_abstraction(gw_m3, par[24], 0), z["d1/wu_daily_nug"])
    assert (z["d1/wu_daily_nus"] != 0).sum() > 100


def test_wateruse_update_net_abstraction_gw_bit_exact():
    """oracle/water_use.py::update_net_abstraction_gw against the compiled reference's updateNetAbstractionGW on 1000 seeded
    cells that reach every branch (return flows reduced / reintroduced, no irrigation withdrawal, dead band, ratio limit)"""
    from oracle import water_use as wu
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_ng1000_wateruse.npz"))
    vin, ref = z["unit/in"], z["unit/out"]
    got = wu.update_net_abstraction_gw(*vin)
    for k, name in enumerate(("NAg", "dailyRemainingUse", "unsatisfiedNAsFromIrrig", "unsatisfiedNAsFromOtherSectors", "reducedReturnFlow")):
        assert np.array_equal(got[k], ref[k]), name
    assert np.array_equal(ref[5], vin[7])  # G_dailydailyNUg itself is left to the caller
    assert (got[0] != vin[7]).sum() > 300 and (got[1] == 0).all() or (np.abs(got[1]) <= 1e-12).sum() > 0


def _wateruse_oracle(golden, oracle_lib):
    """oracle on the 1000-cell golden world with net abstractions (subtract_use 2), started like the reference"""
    from oracle import synth_world as sw, water_use as wu, wg_init
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_ng1000_wateruse.npz"))
    ng = int(z["ng"])
    w = sw.build_world(ng)
    ini = wg_init.derive(w)
    o = oracle_lib.Oracle(ng)
    for k, v in ini.items():
        if not k.startswith("_") and o.has(k):
            o.set(k, v)
    o.set("wu_frgi", z["input/G_FRACTRETURNGW_IRRIG.UNF0"].astype(np.float64))
    o._L.wgo_set_subtract_use(o._c, 2)
    files = {k[6:]: z[k] for k in z.files if k.startswith("input/")}
    return z, w, ini, o, files, wu


def test_wateruse_oracle_bit_exact_vs_reference_golden(golden, oracle_lib):
    """the restatement of the use satisfaction inside routing() (net abstraction from groundwater with
    updateNetAbstractionGW in front of each groundwater balance; surface-water use taken from global lake, reservoir - incl. the
    release rule of irrigation reservoirs -, river and finally the local lake; unsatisfied use and return-flow bookkeeping) is
    BIT-identical to the compiled reference over January and February: all storages, fluxes and the water-use arrays"""
    from oracle import synth_world as sw
    z, w, ini, o, files, wu = _wateruse_oracle(golden, oracle_lib)
    par = np.asarray(ini["params"]).reshape(26, -1)
    n = 0
    for sd in range(1, 60):
        doy, mon, dom = oracle_lib.calendar(sd)
        if dom == 1:
            o.set_forcing_month(sw.forcing_month(w, 1901, mon + 1))
            for k, v in wu.month_inputs(files, par, mon).items():
                o.set(k, v)
        o.step_day(doy, mon, dom)
        if sd in (1, 2, 31, 59):
            for key in z.files:
                if key.startswith(f"d{sd}/"):
                    name = key.split("/", 1)[1]
                    if o.has(name) and name not in ("status_laf_next",):
                        assert np.array_equal(z[key], o.field(name)), f"day {sd} {name}"
                        n += 1
    assert n > 150
    assert (o.field("wu_total_unsatisfied") > 0).sum() > 50 and (o.field("wu_red_rf") != 0).any()
