"""GPU (-m gpu): the timed configuration itself, the station record over several calls, device-side set-up of
large ensembles, the ensemble-moments kernel and its NCCL all-reduce, long single calls against the oracle."""
import os
import socket

import numpy as np
import pytest

from tests.util import ParityReport

pytestmark = pytest.mark.gpu


def _full_world():
    from oracle import synth_world as sw, wg_init
    w = sw.build_world(67420)
    return w, wg_init.derive(w)


def _year_model(w, ini, forcing, **kw):
    import watergap2_b200 as wg
    topo = ini["_topology"]
    m = wg.Model(w.ng, **kw)
    m.set_topology(topo["rout_order"], topo["outflow_cell"], cell_class=wg.cell_classes(ini))
    m.load(ini)
    m.forcing_reserve(365)
    slot = 0
    for mon, f in enumerate(forcing):
        nd = (31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31)[mon]
        m.set_forcing(slot, nd, f["P"], f["T"], f["SW"], f["LW"])
        slot += nd
    return m


def test_year_call_bit_equal_to_daily_calls_full_size():
    """THE TIMED PATH of bench.py: one wgk_step_days(365) call at 67 420 cells - a graph of ~41 600 (day, level) tasks, the 32-slot
    discharge ring wrapping 11 times - must equal, bit for bit, 365 one-day calls and the same call with plain launches
    (use_graph = 0), in every state / flux field, the snow bands and the daily station record."""
    from oracle import synth_world as sw, wg_init
    w, ini = _full_world()
    forcing = [sw.forcing_month(w, 1901, mon) for mon in range(1, 13)]
    names = wg_init.STATE_FIELDS + wg_init.FLUX_FIELDS + ["discharge", "snow_bands"]
    stations = np.argsort(-w.acc)[:50].astype(np.int32)
    out = []
    for mode in ("year_graph", "daily_calls", "year_plain"):
        m = _year_model(w, ini, forcing, use_graph=0 if mode == "year_plain" else 1)
        m.record_cells(stations, 365)
        if mode == "daily_calls":
            from oracle import wgo
            for sd in range(1, 366):
                doy, mon, dom = wgo.calendar(sd)
                m.step_days(doy, mon, dom, sd - 1, 1)
        else:
            m.step_days(1, 0, 1, 0, 365)
        m.synchronize()
        out.append(({k: m.get(k) for k in names}, m.get_record(365)))
        m.close()
    assert np.abs(out[0][1]).sum() > 0 and np.isfinite(out[0][1]).all()
    for other in out[1:]:
        assert np.array_equal(out[0][1], other[1])
        for k in names:
            assert np.array_equal(out[0][0][k], other[0][k]), k


def test_record_accumulates_over_calls(world3000):
    """the station record is indexed by a running day counter, so a multi-year calibration run (several
    calls) finds every year in it; a call that does not fit restarts the record; reading more rows than recorded fails"""
    from oracle import synth_world as sw, wg_init
    import watergap2_b200 as wg
    ini = wg_init.derive(world3000)
    topo = ini["_topology"]
    f = sw.forcing_month(world3000, 1901, 1)
    cells = np.arange(5, world3000.ng, 61, dtype=np.int32)

    def model():
        m = wg.Model(world3000.ng, nmember=2)
        m.set_topology(topo["rout_order"], topo["outflow_cell"], cell_class=wg.cell_classes(ini))
        m.load(ini)
        m.forcing_reserve(31)
        m.set_forcing(0, 31, f["P"], f["T"], f["SW"], f["LW"])
        return m

    a = model()
    a.record_cells(cells, 93)
    for _ in range(3):  # three 31-day "years" in three calls
        a.step_days(1, 0, 1, 0, 31)
    rec = a.get_record(93, 1)
    b = model()
    b.record_cells(cells, 31)
    for y in range(3):  # the same run with a one-year record: every call restarts it
        b.step_days(1, 0, 1, 0, 31)
        assert np.array_equal(b.get_record(31, 1), rec[31 * y:31 * (y + 1)]), y
    assert np.abs(rec[62:]).sum() > 0
    with pytest.raises(wg.WgkError):
        b.get_record(32, 0)  # more than max_days
    b.record_rewind()
    b.step_days(1, 0, 1, 0, 5)
    with pytest.raises(wg.WgkError, match="holds 5 days"):
        b.get_record(6, 0)
    from watergap2_b200 import calibration as cal
    crit = cal.sweep_criteria(a, np.full((3, cells.size), 1.0, np.float32), 3, days_per_year=31)
    assert len(crit) == 2 and crit[0][0]["years"] == 3


def test_clone_and_fill_equal_host_upload(world3000):
    """wgk_copy_index / wgk_fill_field (set-up of 1024 parameter sets / 256 members on the device) give the same bits as
    uploading every member and parameter set from the host"""
    from oracle import synth_world as sw, wg_init
    import watergap2_b200 as wg
    w = world3000
    ini = wg_init.derive(w)
    topo = ini["_topology"]
    f = sw.forcing_month(w, 1901, 1)
    pb = np.array(ini["params"], np.float64).reshape(26, -1).copy()
    pb[0, :], pb[1, :], pb[7, :], pb[15, :] = 2.75, 0.9, 0.02, 0.5
    ini2 = dict(ini)
    ini2["params"], ini2["gamma_hbv"], ini2["cfa"] = pb, pb[0].copy(), pb[1].copy()
    out = []
    for device_side in (False, True):
        m = wg.Model(w.ng, nmember=3, npset=2)
        m.set_topology(topo["rout_order"], topo["outflow_cell"], cell_class=wg.cell_classes(ini))
        if device_side:
            m.load(ini, member=0, pset=0)
            m.copy_pset(0, 1)
            m.copy_member(0, 1)
            m.copy_member(0, 2)
            for name, v in (("gamma_hbv", 2.75), ("cfa", 0.9), ("p_swoutf", 0.02), ("p_snowfz", 0.5)):
                m.fill(name, v, index=1)
        else:
            m.load(ini, pset=0)
            m.load(ini2, pset=1, only={k for k in ini2 if m.has_field(k) and m.field_info(k)[2] == 1})
        m.set_member_pset(0, 0)
        m.set_member_pset(1, 1)
        m.set_member_pset(2, 1)
        m.forcing_reserve(31)
        m.set_forcing(0, 31, f["P"], f["T"], f["SW"], f["LW"])
        m.step_days(1, 0, 1, 0, 9)
        out.append([{k: m.get(k, mem) for k in wg_init.STATE_FIELDS + ["discharge", "snow_bands"]} for mem in range(3)])
        with pytest.raises(wg.WgkError):
            m.fill("smax", 1.0)  # not an f64 field
    for mem in range(3):
        for k in out[0][mem]:
            assert np.array_equal(out[0][mem][k], out[1][mem][k]), (mem, k)
    assert not np.array_equal(out[0][0]["soil"], out[0][1]["soil"])
    assert np.array_equal(out[0][1]["soil"], out[0][2]["soil"])


def _ensemble_model(w, ini, nmember, device=0, seed=3, member0=0):
    """members with their own perturbed forcing (global member index seeds the perturbation)"""
    from oracle import synth_world as sw
    import watergap2_b200 as wg
    topo = ini["_topology"]
    base = sw.forcing_month(w, 1901, 1)
    m = wg.Model(w.ng, nmember=nmember, device=device)
    m.set_topology(topo["rout_order"], topo["outflow_cell"], cell_class=wg.cell_classes(ini))
    m.load(ini)
    m.forcing_reserve(31, per_member=True)
    for k in range(nmember):
        rng = np.random.default_rng([seed, member0 + k])
        P = (base["P"] * np.exp(rng.normal(0., 0.1, base["P"].shape))).astype(np.float32)
        T = (base["T"] + rng.normal(0., 1., base["T"].shape)).astype(np.float32)
        m.set_forcing(0, 31, P, T, base["SW"], base["LW"], member=k)
        m.synchronize()
    return m


def test_moments_kernel_equals_sequential_host_sum(world3000):
    """k_ensemble_moments: sum and sum of squares of the extract_sub_ state vector over the members, bit-equal to adding
    the wgk_state_vector values member by member on the host; all cells and a region; mean / variance of k_moments_finish"""
    from oracle import wg_init
    from watergap2_b200.ensemble import device_tensor, ensemble_state_moments
    w = world3000
    ini = wg_init.derive(w)
    nm = 5
    m = _ensemble_model(w, ini, nm)
    m.month_begin()
    m.step_days(1, 0, 1, 0, 12)
    for kind in ("month", "lastday"):
        for cells in (None, np.arange(7, w.ng, 13, dtype=np.int32)):
            ps, pq, n = m.ensemble_moments(kind, cells)
            m.synchronize()
            got = device_tensor(ps, (2, n, 10), 0).cpu().numpy()
            cl = np.arange(w.ng, dtype=np.int32) if cells is None else cells
            s, ss = np.zeros((n, 10)), np.zeros((n, 10))
            for k in range(nm):
                v = m.state_vector(cl, kind, member=k)
                s += v
                ss += v * v
            assert np.array_equal(got[0], s) and np.array_equal(got[1], ss), (kind, cells is None)
    mean, var = ensemble_state_moments(m, nm, "month")
    vs = np.stack([m.state_vector(np.arange(w.ng, dtype=np.int32), "month", member=k) for k in range(nm)])
    assert np.allclose(mean, vs.mean(0), rtol=1e-13, atol=1e-300)
    assert np.allclose(var, vs.var(0), rtol=1e-6, atol=1e-9 * np.abs(vs).max())
    assert var[:, 2].max() > 0  # members really differ (soil)


def _nccl_worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from oracle import synth_world as sw, wg_init
    from watergap2_b200.ensemble import ensemble_state_moments, shard_members
    w = sw.build_world(3000)
    ini = wg_init.derive(w)
    total = 7
    first, count = shard_members(total, world, rank)
    m = _ensemble_model(w, ini, count, device=rank, member0=first)
    m.month_begin()
    m.step_days(1, 0, 1, 0, 10)
    t = {}
    mean, var = ensemble_state_moments(m, total, "month", timing=t)
    cells = np.arange(w.ng, dtype=np.int32)
    mine = torch.from_numpy(np.stack([m.state_vector(cells, "month", member=k) for k in range(count)])).cuda()
    pad = torch.zeros((4,) + tuple(mine.shape[1:]), dtype=torch.float64, device="cuda")
    pad[:count] = mine
    allv = torch.empty((world * 4,) + tuple(mine.shape[1:]), dtype=torch.float64, device="cuda")
    dist.all_gather_into_tensor(allv, pad)
    if rank == 0:
        a = allv.cpu().numpy().reshape(world, 4, w.ng, 10)
        vs = np.concatenate([a[r, :shard_members(total, world, r)[1]] for r in range(world)])
        torch.save({"mean": mean, "var": var, "ref_mean": vs.mean(0), "ref_var": vs.var(0), "n": vs.shape[0], "t": t}, out)
    dist.destroy_process_group()


def test_ensemble_moments_over_nccl_world_size_2(tmp_path):
    """SURVEY 8e config 4 on hardware: members sharded over 2 GPUs, k_ensemble_moments + ONE NCCL all-reduce on the
    context's stream + k_moments_finish against the numpy statistic of all members (needs 2 GPUs: `gpurun --gpus 2`)"""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "mom.pt")
    mp.spawn(_nccl_worker, args=(2, port, out), nprocs=2, join=True)
    r = torch.load(out, weights_only=False)
    assert r["n"] == 7
    assert np.allclose(r["mean"], r["ref_mean"], rtol=1e-13, atol=1e-300)
    assert np.allclose(r["var"], r["ref_var"], rtol=1e-6, atol=1e-9 * np.abs(r["ref_mean"]).max())
    assert r["t"]["allreduce_bytes"] == 2 * 3000 * 10 * 8


def test_long_single_calls_vs_oracle(world3000):
    """calls longer than the 32-day discharge ring against the oracle: three single calls of 40 days, both sides
    re-synchronised to the oracle's state at the call boundaries (free run inside a call: policy of DESIGN.md 6)"""
    from oracle import synth_world as sw, wg_init, wgo
    import watergap2_b200 as wg
    w = world3000
    ini = wg_init.derive(w)
    topo = ini["_topology"]
    o = wgo.Oracle(w.ng)
    for k, v in ini.items():
        if not k.startswith("_") and o.has(k):
            o.set(k, v)
    m = wg.Model(w.ng)
    m.set_topology(topo["rout_order"], topo["outflow_cell"], cell_class=wg.cell_classes(ini))
    m.load(ini)
    m.forcing_reserve(365)
    forcing = [sw.forcing_month(w, 1901, mon) for mon in range(1, 6)]
    slot = 0
    for mon, f in enumerate(forcing):
        nd = (31, 28, 31, 30, 31)[mon]
        m.set_forcing(slot, nd, f["P"], f["T"], f["SW"], f["LW"])
        slot += nd
    names = wg_init.STATE_FIELDS + wg_init.FLUX_FIELDS
    sd = 1
    for block in range(3):
        if block:
            for name in wg_init.STATE_FIELDS + ["storage_transfer"]:
                m.set(name, o.field(name))
        doy, mon, dom = wgo.calendar(sd)
        m.step_days(doy, mon, dom, sd - 1, 40)
        for k in range(40):
            doy, mon, dom = wgo.calendar(sd + k)
            if dom == 1:
                o.set_forcing_month(forcing[mon])
            o.step_day(doy, mon, dom)
        sd += 40
        rep = ParityReport()
        for name in names:
            rep.add(name, o.field(name), m.get(name), tag=sd - 1)
        cells = {f[2] for f in rep.flips}
        assert len(cells) <= max(2, w.ng // 500) and rep.worst < 1e-6, (rep.summary(), sorted(rep.flips, key=lambda f: -f[5])[:10])


def test_member_minor_layout_is_bit_identical(world3000, monkeypatch):
    """lane = member: the member-minor layout [band][cell][member] (a warp = 32 members of one cell, chosen automatically for
    large ensembles / parameter sweeps) against the cell-minor layout: 40 members (padded to 64 lanes) with 40 different
    parameter sets and per-member forcing, 14 days; every state / flux field of several members, the snow bands, the station
    record, the monthly state vector, the ensemble moments and the field round trip must be the same bits - with the whole-day
    schedule and with the (day, level) wavefront graph on the member-minor layout (the schedule of 32 - 64 members per GPU)"""
    from oracle import synth_world as sw, wg_init
    import watergap2_b200 as wg
    from watergap2_b200.ensemble import device_tensor
    w = world3000
    ini = wg_init.derive(w)
    topo = ini["_topology"]
    base = sw.forcing_month(w, 1901, 1)
    nm = 40
    names = wg_init.STATE_FIELDS + wg_init.FLUX_FIELDS + ["discharge", "snow_bands"]
    rng = np.random.default_rng(5)
    forc = []
    for k in range(nm):
        forc.append(((base["P"] * np.exp(rng.normal(0., 0.1, base["P"].shape))).astype(np.float32),
                     (base["T"] + rng.normal(0., 1.5, base["T"].shape)).astype(np.float32)))
    cells = np.arange(3, w.ng, 17, dtype=np.int32)
    out = []
    for layout, sched in (("cells", "wholeday"), ("members", "wholeday"), ("members", "wavefront")):
        monkeypatch.setenv("WGK_LAYOUT", layout)
        monkeypatch.setenv("WGK_DAY_SCHEDULE", sched)
        monkeypatch.setenv("WGK_VERTICAL_FORM", "cells")
        m = wg.Model(w.ng, nmember=nm, npset=nm)
        assert m.layout == layout
        m.set_topology(topo["rout_order"], topo["outflow_cell"], cell_class=wg.cell_classes(ini))
        m.load(ini, member=0, pset=0)
        for k in range(1, nm):
            m.copy_pset(0, k)
            m.copy_member(0, k)
            m.fill("gamma_hbv", 0.5 + 0.1 * k, index=k)
            m.fill("p_snowfz", -1.0 + 0.05 * k, index=k)
            m.fill("p_gwoutf", 0.005 + 0.001 * k, index=k)
        soil7 = ini["soil"] + 3.25
        m.set("soil", soil7, 7)  # host upload of one member's field in either layout
        assert np.array_equal(m.get("soil", 7), soil7) and np.array_equal(m.get("soil", 6), ini["soil"])
        assert np.array_equal(m.get("gamma_hbv", 3), np.full(w.ng, 0.5 + 0.1 * 3))
        m.forcing_reserve(31, per_member=True)
        for k in range(nm):
            m.set_forcing(0, 31, forc[k][0], forc[k][1], base["SW"], base["LW"], member=k)
            m.synchronize()
        m.record_cells(cells[:20], 31)
        m.step_days(1, 0, 1, 0, 4)
        m.month_begin()
        m.step_days(5, 0, 5, 4, 10)
        ps, pq, n = m.ensemble_moments("month")
        m.synchronize()
        mom = device_tensor(ps, (2, n, 10), 0).cpu().numpy()
        out.append(({(k, mem): m.get(k, mem) for k in names for mem in (0, 7, 31, 32, 39)}, m.get_record(14, 33),
                    m.state_vector(cells, "month", member=38), mom, m.total_storage_km3(39)))
        m.close()
    assert np.abs(out[0][1]).sum() > 0
    for other in out[1:]:
        assert np.array_equal(out[0][1], other[1])
        assert np.array_equal(out[0][2], other[2])
        assert np.array_equal(out[0][3], other[3])
        assert out[0][4] == other[4]
        for k in out[0][0]:
            assert np.array_equal(out[0][0][k], other[0][k]), k
    assert not np.array_equal(out[0][0][("soil", 0)], out[0][0][("soil", 39)])


def test_water_use_vs_reference_golden(golden, monkeypatch):
    """SURVEY 8f-4 on the GPU: net abstractions from surface water and groundwater (subtract_use 2) on the 1000-cell golden world,
    January and February, against what the COMPILED REFERENCE held in memory (tests/golden/ref_ng1000_wateruse.npz): storages,
    fluxes and the water-use bookkeeping (unsatisfied use, adapted groundwater abstraction, reduced return flows, actual use).
    Days 1 and 2 must hold 1e-10 without exception; days 31 and 59 under the free-run policy, listed.  Both schedules."""
    from oracle import synth_world as sw, water_use as wu, wg_init
    import watergap2_b200 as wg
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_ng1000_wateruse.npz"))
    ng = int(z["ng"])
    w = sw.build_world(ng)
    ini = wg_init.derive(w)
    topo = ini["_topology"]
    par = np.asarray(ini["params"]).reshape(26, -1)
    files = {k[6:]: z[k] for k in z.files if k.startswith("input/")}
    wu_names = ["wu_total_unsatisfied", "wu_daily_remaining", "wu_daily_nug", "wu_actual_use", "wu_uns_irr", "wu_uns_oth", "wu_red_rf", "wu_wusi", "wu_cusi"]
    for sched in ("wavefront", "wholeday"):
        monkeypatch.setenv("WGK_DAY_SCHEDULE", sched)
        m = wg.Model(ng, subtract_use=2)
        m.set_topology(topo["rout_order"], topo["outflow_cell"], cell_class=wg.cell_classes(ini))
        m.load(ini)
        m.set("wu_frgi", z["input/G_FRACTRETURNGW_IRRIG.UNF0"].astype(np.float64))
        m.forcing_reserve(31)
        sd = 1
        nchk = 0
        for mon, ndays in ((0, 31), (1, 28)):
            f = sw.forcing_month(w, 1901, mon + 1)
            m.set_forcing(0, 31, f["P"], f["T"], f["SW"], f["LW"])
            for k, v in wu.month_inputs(files, par, mon).items():
                m.set(k, v)
            done = 0
            for stop in ((1, 2, 31) if mon == 0 else (28,)):
                n = stop - done
                doy = sd
                m.step_days(doy, mon, done + 1, done, n)
                sd += n
                done = stop
                day = sd - 1
                rep = ParityReport()
                for key in z.files:
                    if key.startswith(f"d{day}/"):
                        name = key.split("/", 1)[1]
                        if m.has_field(name) and name != "status_laf_next":
                            rep.add(name, z[key], m.get(name), tag=day)
                            nchk += 1
                print(sched, "day", day, rep.summary(), sorted(rep.flips, key=lambda f: -f[5])[:6])
                if day <= 2:
                    assert not rep.flips, rep.flips[:10]
                else:  # free run on the small golden world: the handful of cells at the evaporation-limited river threshold
                    cells = {f[2] if f[1] != "snow_bands" else f[2] // 101 for f in rep.flips}
                    assert len(cells) <= 12 and rep.worst < 1e-4, (rep.summary(), sorted(rep.flips, key=lambda f: -f[5])[:10])
        assert nchk > 150
        assert (m.get("wu_total_unsatisfied") > 0).sum() > 50 and (m.get("wu_red_rf") != 0).any() and (m.get("gw") < 0).any()
        m.close()
    with pytest.raises(wg.WgkError, match="subtract_use"):
        wg.Model(ng).get("wu_red_rf")  # the water-use arrays exist only with water use


def test_water_use_one_step_parity_59_days(golden, oracle_lib):
    """water use, one step at a time: every day both sides start from the ORACLE's state (which is bit-identical to the compiled
    reference with water use, tests/test_oracle_golden.py), take one step and are compared - the kernels on every day of January
    and February without the growth of a free run: the one-step policy of tests/util.py, water-use bookkeeping included"""
    from oracle import synth_world as sw, water_use as wu, wg_init
    from tests.util import ONE_STEP_MAX_REL, ONE_STEP_PPM
    import watergap2_b200 as wg
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_ng1000_wateruse.npz"))
    ng = int(z["ng"])
    w = sw.build_world(ng)
    ini = wg_init.derive(w)
    topo = ini["_topology"]
    par = np.asarray(ini["params"]).reshape(26, -1)
    files = {k[6:]: z[k] for k in z.files if k.startswith("input/")}
    o = oracle_lib.Oracle(ng)
    for k, v in ini.items():
        if not k.startswith("_") and o.has(k):
            o.set(k, v)
    o.set("wu_frgi", z["input/G_FRACTRETURNGW_IRRIG.UNF0"].astype(np.float64))
    o._L.wgo_set_subtract_use(o._c, 2)
    m = wg.Model(ng, subtract_use=2)
    m.set_topology(topo["rout_order"], topo["outflow_cell"], cell_class=wg.cell_classes(ini))
    m.load(ini)
    m.set("wu_frgi", o.field("wu_frgi"))
    m.forcing_reserve(31)
    wu_state = ["wu_total_unsatisfied", "wu_daily_remaining", "wu_uns_irr", "wu_uns_oth", "wu_red_rf", "wu_wusi", "wu_cusi", "wu_actual_use"]
    rep = ParityReport()
    for sd in range(1, 60):
        doy, mon, dom = oracle_lib.calendar(sd)
        if dom == 1:
            f = sw.forcing_month(w, 1901, mon + 1)
            o.set_forcing_month(f)
            m.set_forcing(0, 31, f["P"], f["T"], f["SW"], f["LW"])
            for k, v in wu.month_inputs(files, par, mon).items():
                o.set(k, v)
                m.set(k, v)
        if sd > 1:
            for name in wg_init.STATE_FIELDS + ["storage_transfer"] + wu_state:
                m.set(name, o.field(name))
        o.step_day(doy, mon, dom)
        m.step_days(doy, mon, dom, dom - 1, 1)
        for name in wg_init.STATE_FIELDS + wg_init.FLUX_FIELDS + wu_state + ["wu_daily_nug"]:
            rep.add(name, o.field(name), m.get(name), tag=sd)
    print("water use, one step:", rep.summary(), sorted(rep.flips, key=lambda f: -f[5])[:8])
    rep.check(ONE_STEP_PPM, ONE_STEP_MAX_REL, min_allowed=2, what="water use, one step, 59 days")
    assert (o.field("wu_total_unsatisfied") > 0).sum() > 50


def test_fused_level_tasks_are_bit_identical(world3000, monkeypatch):
    """task forms of the wavefront graph (WGK_LEVEL_TASKS): vertical + river part of a level in one kernel with a programmatic edge
    from the upstream level ("fused", k_level_day with griddepcontrol.wait - the default of small problems - with every level a
    fused task or with the narrow levels as tail chunks), with a full edge ("fusedfull"), or for the headwater level only
    ("fused0"), against the two-kernel tasks ("split"): 45 days in one call, every state / flux field, the snow bands and the
    station record the same bits"""
    from oracle import synth_world as sw, wg_init
    import watergap2_b200 as wg
    w = world3000
    ini = wg_init.derive(w)
    topo = ini["_topology"]
    names = wg_init.STATE_FIELDS + wg_init.FLUX_FIELDS + ["discharge", "snow_bands"]
    cells = np.arange(3, w.ng, 17, dtype=np.int32)[:20]
    monkeypatch.setenv("WGK_VERTICAL_FORM", "cells")
    monkeypatch.setenv("WGK_TAIL_THRESHOLD", "16")  # several wide levels on the small world
    out = []
    for mode, wave_tail in (("split", ""), ("fused", "0"), ("fused", "16"), ("fusedfull", "0"), ("fused0", "")):
        monkeypatch.setenv("WGK_LEVEL_TASKS", mode)
        monkeypatch.setenv("WGK_WAVE_TAIL_THRESHOLD", wave_tail) if wave_tail else monkeypatch.delenv("WGK_WAVE_TAIL_THRESHOLD", raising=False)
        m = wg.Model(w.ng)
        assert (m.schedule & 2 != 0) == (mode != "split")
        m.set_topology(topo["rout_order"], topo["outflow_cell"], cell_class=wg.cell_classes(ini))
        m.load(ini)
        m.forcing_reserve(59)
        slot = 0
        for mon, nd in ((1, 31), (2, 28)):
            f = sw.forcing_month(w, 1901, mon)
            m.set_forcing(slot, nd, f["P"], f["T"], f["SW"], f["LW"])
            slot += nd
        m.record_cells(cells, 45)
        m.step_days(1, 0, 1, 0, 45)
        m.synchronize()
        out.append(({k: m.get(k) for k in names}, m.get_record(45)))
        m.close()
    for st, rec in out[1:]:
        for k in names:
            assert np.array_equal(st[k], out[0][0][k]), k
        assert np.array_equal(rec, out[0][1])


@pytest.mark.gpu
def test_default_task_form_follows_problem_size(monkeypatch):
    """wgk_create picks the fused (day, level) tasks for latency-bound problems (cell-minor layout, below 800 000 cell-members)
    and the two-kernel tasks beyond, and says so in wgk_schedule (bit 2); water use keeps the two-kernel tasks"""
    import watergap2_b200 as wg
    for k in ("WGK_LEVEL_TASKS", "WGK_DAY_SCHEDULE", "WGK_LAYOUT"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setenv("WGK_VERTICAL_FORM", "cells")
    for ncell, nmember, use, fused in ((67420, 1, 0, True), (67420, 8, 0, True), (67420, 16, 0, False), (900000, 1, 0, False), (67420, 1, 2, False)):
        m = wg.Model(ncell, nmember=nmember, subtract_use=use)
        assert (m.schedule & 1) == 0, (ncell, nmember)
        assert ((m.schedule & 2) != 0) == fused, (ncell, nmember, use, m.schedule)
        m.close()
