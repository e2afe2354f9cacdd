#!/usr/bin/env python3
"""bench.py — simulated cell-days per second of the WaterGAP2 daily hot path on B200.

  python bench.py --gpus N --steps K --warmup W [--impl wgk|reference] [--members M] [--legs sweep,enkf,basins|none]

Headline workload (BASELINE.json configs[1]): the 0.5 degree global synthetic grid (67 420 cells, seed
20240607), daily time step with routing and 100 elevation-band snow.  One STEP is one
simulated model year (365 days); the default K = 30 timed steps is the 30-year run of the
config.  Per GPU one member (a single model run) unless --members is given; at N > 1 every
rank runs its own member(s) (replicas: a single 0.5 degree member does not shard).

value   = cells x 365 x members x N x K / max-over-ranks(device time of the K steps), state and a
          full year of forcing resident in HBM (forcing slots are cycled on the device).
e2e     = the same metric through the C ABI with HOST buffers: every step the year's forcing
          (12 x 4 grids [ncell][31] float32, the reference's .31 layout) is copied from pinned
          host memory and packed on the device, the year is stepped, and the daily discharge of
          50 station cells plus the last day's per-cell discharge are read back to the host.
roofline, cpu_baseline: see DESIGN.md 5.

sharded = the three configurations of BASELINE.json that DO shard over the GPUs, measured in the same process at
          --gpus N after the headline (each with a parity sample against the CPU oracle, outside its timed region):
  sweep_1024   configs[2]: 1024 calibration parameter sets, contiguous blocks per rank, one simulated month per
               step, the annual station runoff of all sets all-gathered over NCCL inside the timed region;
  enkf_256     configs[3]: 256 ensemble members with per-member forcing, per step one simulated month, the
               moments kernel over the [ncell x 10] state vector and ONE NCCL all-reduce of sum | sumsq;
  basins_5arcmin  configs[4]: a 2 157 440-cell grid split by whole drainage basin (strong scaling).

--impl reference times the reference's own CPU implementation (oracle/_ref harness: the
unmodified daily.cpp/routing.cpp driven through the replayed day loop of integrateWGHM.cpp)
on the host cores, each step a bounded sample of the same workload.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NG = 67420
NDAYS = [31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]
BYTES_VERTICAL, BYTES_ROUTING = 2099, 617  # algorithmic bytes per cell-day, SURVEY.md 8(d)
BYTES_LOCAL_ROUTING = 300                  # share of the 617 B that the cell-parallel local routing inside k_cells_pre* touches
FLOP_PER_CELL_DAY = 2700 + 900             # FP64 operations per cell-day, SURVEY.md 8(d) (vertical + routing, transcendental = 50)
METRIC = "simulated cell-days/sec, 0.5deg global grid"
UNIT = "cell-days/s"
SWEEP_FIELDS = {0: "gamma_hbv", 1: "cfa", 7: "p_swoutf", 22: "p_gwoutf", 15: "p_snowfz", 16: "p_snowmt", 4: "p_rivrgh",
                8: "p_evaredex", 9: "p_netrad", 17: "p_degday", 25: "p_prec"}  # eCalibParam -> per-cell f64 device array


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def build_inputs():
    from oracle import synth_world as sw, wg_init
    w = sw.build_world(NG)
    ini = wg_init.derive(w)
    return w, ini


def year_forcing(w):
    from oracle import synth_world as sw
    return [sw.forcing_month(w, 1901, m) for m in range(1, 13)]


def host_model(w, device):
    """the headline model through the product's own HOST LAYER: the synthetic world is written as the reference's input files
    (UNF grids, OPTIONS.DAT, parameter JSON, config.txt) and libwghost.so builds the routing files, runs the init sequence of
    integrate_wghm_ and pushes statics / parameters / start state to the device (wg_host_create_context)"""
    import shutil
    import watergap2_b200 as wg
    from oracle import synth_world as sw
    tmp = tempfile.mkdtemp(prefix="wg_bench_world_")
    try:
        sw.write_world(w, tmp, (1901, 1901), (1, 1), grid_store=0, daily_discharge=False)
        m = wg.Model.from_config(os.path.join(tmp, "config.txt"), w.ng, device)
        t0 = time.perf_counter()
        secs = ctypes_double()
        err = ctypes_buf(1024)
        nd = wg.host_lib().wg_host_integrate(os.fsencode(os.path.join(tmp, "config.txt")), w.ng, device, secs_ref(secs), err, 1024)
        classes = None
        if nd > 0:
            classes = {"value": w.ng * nd / secs.value, "unit": UNIT, "days": int(nd), "seconds_day_loop": round(secs.value, 4),
                       "wall_s_with_init_and_files": round(time.perf_counter() - t0, 2),
                       "what": "January 1901 through the drop-in C++ classes (libwghost.so, wg_host_integrate): per-cell calcNewDay shim, "
                               "routingClass::routing per day with one packed device-to-host copy of the day's WghmStateFile entry, "
                               "updateLandAreaFrac, month-end rescale; reference-format files in, three checkpoint files out"}
        else:
            classes = {"error": err.value.decode()}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return m, classes


def ctypes_double():
    import ctypes
    return ctypes.c_double()


def ctypes_buf(n):
    import ctypes
    return ctypes.create_string_buffer(n)


def secs_ref(x):
    import ctypes
    return ctypes.byref(x)


def make_model(w, ini, members, device, npset=1):
    import watergap2_b200 as wg
    m = wg.Model(w.ng, nmember=members, npset=npset, device=device)
    topo = ini["_topology"]
    m.set_topology(topo["rout_order"], topo["outflow_cell"], cell_class=wg.cell_classes(ini))
    if members == 1 and npset == 1:
        m.load(ini)
    else:  # one upload, then device-to-device replication (wgk_copy_index)
        m.load(ini, member=0, pset=0)
        for k in range(1, npset):
            m.copy_pset(0, k)
        for k in range(1, members):
            m.copy_member(0, k)
    return m


def upload_year(m, forcing, slot0=0, reserve=True):
    if reserve:
        m.forcing_reserve(730)  # two years of slots: the e2e loop uploads one year while the other is stepped
    slot = slot0
    for mon in range(12):
        f = forcing[mon]
        m.set_forcing(slot, NDAYS[mon], f["P"], f["T"], f["SW"], f["LW"])
        slot += NDAYS[mon]
    if reserve:
        m.synchronize()


class Dist:
    """one process per GPU; NCCL for the barrier, the max over ranks and the data-path collectives of the sharded legs"""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device - the wgk path has no CPU fallback")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, x, op="max"):
        t = self.torch.tensor([float(x)], device="cuda", dtype=self.torch.float64)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return float(t.item())

    def gather(self, x):
        t = self.torch.tensor([float(x)], device="cuda", dtype=self.torch.float64)
        if self.world == 1:
            return [float(x)]
        out = self.torch.empty(self.world, device="cuda", dtype=self.torch.float64)
        self.dist.all_gather_into_tensor(out, t)
        return [float(v) for v in out.cpu()]

    def timed(self, m, fn, steps):
        """barrier + synchronize, `steps` calls of fn() timed by CUDA events on the context's stream, barrier +
        synchronize; -> max over ranks of the device time in ms"""
        torch = self.torch
        stream = torch.cuda.ExternalStream(m.stream, device=self.local)
        m.synchronize()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for k in range(steps):
            fn(k)
        e1.record(stream)
        m.synchronize()
        self.barrier()
        return self.reduce(e0.elapsed_time(e1))

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def oracle_for(ini):
    from oracle import wgo
    o = wgo.Oracle(int(np.asarray(ini["area"]).size))
    for k, v in ini.items():
        if not k.startswith("_") and o.has(k):
            o.set(k, v)
    return o


def parity_sample(pairs, names):
    """pairs: [(tag, oracle, getter(name) -> array of the product)] -> summary dict of tests/util.ParityReport"""
    from tests.util import ParityReport
    rep = ParityReport()
    for tag, o, get in pairs:
        for name in names:
            rep.add(name, o.field(name), get(name), tag=tag)
    s = rep.summary()
    return {"parity_checked": True, "worst_rel": s["worst_rel"], "worst_rel_no_floor": s["worst_rel_no_floor"],
            "values_compared": s["values"], "beyond_1e-10": s["beyond_1e-10"],
            "floors": "1e-9 km3 / 1e-6 mm / 1e-6 (tests/util.py)", "worst_at": s["worst_at"]}


PARITY_NAMES = ["canopy", "soil", "snow", "gw", "loc_lake_stor", "loc_wetl_stor", "glo_lake_stor", "glo_wetl_stor", "res_stor",
                "river_stor", "discharge", "land_area_frac", "lai_days", "lai_status", "surface_runoff", "gw_recharge", "snow_bands"]


# ------------------------------------------------------------------------------------------------
# headline: configs[1]
# ------------------------------------------------------------------------------------------------
def run_headline(args, D, w, ini, forcing):
    import torch
    rank, world, local = D.rank, D.world, D.local
    ncell, tiles, shard = w.ng, 1, None
    e2e_classes = None
    if args.workload == "5arcmin":
        m, ncell, forcing, shard = build_basin_model(D, w, ini, forcing, 32, args.members)
        tiles = 32
    elif args.members == 1:
        m, e2e_classes = host_model(w, local)
    else:
        m = make_model(w, ini, args.members, local)
    upload_year(m, forcing)
    stream = torch.cuda.ExternalStream(m.stream, device=local)

    # ---- device-resident throughput ("value") ------------------------------------------------
    for _ in range(args.warmup):
        m.step_days(1, 0, 1, 0, 365)
    m.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = m.kernel_launches
    ms_max = D.timed(m, lambda k: m.step_days(1, 0, 1, 0, 365), args.steps)
    launches = m.kernel_launches - l0
    clocks = sampler.stop() if sampler else None
    total_cells = D.reduce(ncell, "sum") if (world > 1 and tiles > 1) else float(ncell) * world
    cell_days = total_cells * 365 * args.members * args.steps
    value = cell_days / (ms_max / 1e3)
    peak, peak_src = measured_peaks()
    step_bytes = (BYTES_VERTICAL + BYTES_ROUTING) * cell_days / world
    step_ach = step_bytes / (ms_max / 1e3) / 1e9

    # ---- the kernels of the timed schedule, one by one -------------------------------------------
    # (a) every launch of one simulated day timed on its own with CUDA events (plain launches in task order), per
    #     kernel class; (b) for the wavefront graph, the duration of the level-0 tasks INSIDE the running graph from
    #     %globaltimer stamps (first warp start -> last warp end, median over a simulated year)
    nprof = 10
    cls = {"vertical": [0.0, 0], "river_level": [0.0, 0], "tail": [0.0, 0], "other": [0.0, 0]}
    for d in range(nprof):
        p = m.profile_schedule(1 + d, 0, 1 + d, d)
        for k, (t, n) in p.items():
            cls[k][0] += t / nprof
            cls[k][1] = n
    wavefront = cls["other"][1] == 0
    lv = np.bincount(m.levels(), minlength=m.nlevels)
    form = os.environ.get("WGK_VERTICAL_FORM") or "cells"
    fused = wavefront and bool(m.schedule & 2)  # one task per (day, level): k_level_day = vertical + local routing + river + post
    vname = {"cells": "k_cells_pre_tpc", "bands": "k_cells_pre<VCfg<5,4,1>>", "bands2": "k_cells_pre<VCfg<2,5,0>>"}[form] if wavefront else \
            {"cells": "k_vertical_tpc", "bands": "k_vertical<VCfg<5,4,1>>", "bands2": "k_vertical<VCfg<2,5,0>>"}[form]
    if fused and cls["river_level"][1] == 0:
        vname = "k_level_day"
    names = {"vertical": vname, "river_level": "k_river_level" if wavefront else "k_route_level",
             "tail": "k_tail_chunk" if wavefront else "k_route_tail", "other": "k_route_local + k_route_post"}
    day_ms = sum(v[0] for v in cls.values())
    dom_key = max(cls, key=lambda k: cls[k][0])
    bytes_v_day = (BYTES_VERTICAL + (BYTES_ROUTING if vname == "k_level_day" else BYTES_LOCAL_ROUTING if wavefront else 0)) * ncell * args.members
    in_graph = None
    if wavefront and m.nlevels > 0:
        m.stamps(True)
        m.step_days(1, 0, 1, 0, 365)  # first launch of the graph rebuilt with the stamp buffer (upload), not looked at
        m.synchronize()
        m.stamps(True)                # reset
        m.step_days(1, 0, 1, 0, 365)
        m.synchronize()
        st = m.stamps(False, read=True)[:, :, 20:360].astype(np.int64)
        vdur, rdur = st[0, 1] - st[0, 0], st[1, 1] - st[1, 0]
        if vname == "k_level_day":  # one task: first warp into its V part -> last warp out of its R part
            vdur = st[1, 1] - st[0, 0]
        period = st[0, 0][1:] - st[0, 0][:-1]
        n0 = int(lv[0])
        b0 = (BYTES_VERTICAL + (BYTES_ROUTING if vname == "k_level_day" else BYTES_LOCAL_ROUTING)) * n0 * args.members
        in_graph = {"level0_cells": n0, "vertical_task_us": round(float(np.median(vdur)) / 1e3, 2),
                    "river_task_us": round(float(np.median(rdur)) / 1e3, 2), "day_period_us": round(float(np.median(period)) / 1e3, 2),
                    "day_period_mean_us": round(float(np.mean(period)) / 1e3, 2),
                    "vertical_task_bytes": b0, "vertical_task_gbs": round(b0 / (float(np.median(vdur)) * 1e-9) / 1e9, 1),
                    "vertical_task_frac": round(b0 / (float(np.median(vdur)) * 1e-9) / 1e9 / peak, 4),
                    "how": "%globaltimer stamps of the level-0 tasks inside the running 365-day graph (wgk_stamps), median over days 20..360 "
                           "of an extra simulated year outside the timed region; " +
                           ("vertical_task = the whole fused task k_level_day(d, 0) (first warp into its vertical part -> last warp out of "
                            "its river part), river_task = the span of the river parts inside it; the own-cell recurrence "
                            "F(d,l) -> F(d+1,l) of every level bounds a single member" if vname == "k_level_day" else
                            "the level-0 vertical task is the head of the own-cell "
                            "recurrence V(d,0) -> R(d,0) -> V(d+1,0) that bounds a single member")}
    traffic = None
    try:  # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (tools/ncu_summary.py traffic)
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            traffic = json.load(fh).get(vname.split("<")[0], {}).get(str(args.members)) if tiles == 1 else None
    except Exception:
        pass
    fp64_peak = m.fp64_peak_tflops()
    ach_v = bytes_v_day / (cls["vertical"][0] * 1e-3) / 1e9
    roofline = {
        "bound": "hbm", "unit": "GB/s", "peak": peak, "peak_source": peak_src,
        # headline numbers: the WHOLE timed step (every kernel of the simulated year) in algorithmic bytes
        "kernel": f"timed wgk_step_days schedule ({'(day, level) wavefront graph' if wavefront else 'whole-day kernels'}): " +
                  " + ".join(f"{names[k]} ({cls[k][0] / day_ms:.0%} of kernel time)" for k in sorted(cls, key=lambda k: -cls[k][0]) if cls[k][1]),
        "achieved": round(step_ach, 1), "frac": round(step_ach / peak, 4),
        "step_achieved": round(step_ach, 1), "step_frac": round(step_ach / peak, 4),
        "algorithmic_bytes_per_cell_day": BYTES_VERTICAL + BYTES_ROUTING,
        "traffic": traffic,
        "dominant_kernel": {
            "name": names[dom_key], "share_of_kernel_time": round(cls[dom_key][0] / day_ms, 4),
            "launches_per_day": cls["vertical"][1], "isolated_ms_per_day": round(cls["vertical"][0], 5),
            "avg_launch_ms": round(cls["vertical"][0] / max(1, cls["vertical"][1]), 6),
            "algorithmic_bytes_per_day": bytes_v_day, "achieved": round(ach_v, 1), "frac": round(ach_v / peak, 4),
            "traffic_first_launch": traffic,
            "how": "CUDA event pair around every launch of one simulated day of the timed schedule (wgk_profile_schedule, plain "
                   "launches in task order, mean of 10 days); bytes = " + ("(2099 vertical + 617 routing)" if vname == "k_level_day" else
                   "(2099 vertical + 300 local routing)") + " x cells x members",
            "in_graph": in_graph} if dom_key == "vertical" else {"name": names[dom_key], "share_of_kernel_time": round(cls[dom_key][0] / day_ms, 4)},
        "kernel_classes_ms_per_day": {names[k]: {"ms": round(v[0], 5), "launches": v[1]} for k, v in cls.items() if v[1]},
        "fp64": {"peak_tflops_measured": round(fp64_peak, 2), "how": "k_fp64_peak: 8 independent DFMA chains per thread, 8 x 256 threads per SM, best of 3",
                 "step_tflops": round(FLOP_PER_CELL_DAY * cell_days / world / (ms_max / 1e3) / 1e12, 3),
                 "step_frac": round(FLOP_PER_CELL_DAY * cell_days / world / (ms_max / 1e3) / 1e12 / fp64_peak, 4) if fp64_peak > 0 else None,
                 "flop_per_cell_day": FLOP_PER_CELL_DAY},
        "note": "algorithmic bytes count every cell's 100 snow bands (SURVEY 8d); cells without snow and above freezing skip the band loop, "
                "so the measured DRAM traffic is about half of them. A single member is bound by the own-cell day-to-day recurrence "
                "(in_graph.day_period_us), not by HBM; see the sharded legs for the bandwidth-bound regime."}

    # ---- end to end through the C ABI with host buffers ----------------------------------------
    pinned = []
    for mon in range(12):
        pinned.append({k: torch.from_numpy(np.ascontiguousarray(forcing[mon][k])).pin_memory() for k in ("P", "T", "SW", "LW")})
    stations = np.argsort(-w.acc)[:50].astype(np.int32) if tiles == 1 else np.arange(0, ncell, max(1, ncell // 50), dtype=np.int32)[:50]
    m.record_cells(stations, 365)
    out_host = torch.empty(ncell, dtype=torch.float64).pin_memory().numpy()
    e2e_steps = max(1, min(args.steps, 5))

    def upload(year):  # the year's 12 x 4 grids from pinned host memory into the slots of its parity (asynchronous)
        slot = 365 * (year % 2)
        for mon in range(12):
            f = pinned[mon]
            m.set_forcing(slot, NDAYS[mon], f["P"].numpy(), f["T"].numpy(), f["SW"].numpy(), f["LW"].numpy())
            slot += NDAYS[mon]

    def e2e_year(year):
        m.step_days(1, 0, 1, 365 * (year % 2), 365)
        upload(year + 1)  # copy + pack of the next year overlap this year's stepping (own stream, event-ordered)
        res = []
        for mem in range(args.members):
            out_host[:] = m.get("discharge", mem)
            res.append(m.get_record(365, mem))
        return res

    upload(0)
    e2e_year(0)  # warm-up (graph re-instantiation after record_cells)
    m.synchronize()
    D.barrier()
    t0 = time.perf_counter()
    for y in range(1, e2e_steps + 1):
        e2e_year(y)
    m.synchronize()
    D.barrier()
    dt = D.reduce(time.perf_counter() - t0)
    e2e_val = total_cells * 365 * args.members * e2e_steps / dt
    h2d = sum(4 * ncell * 31 * 4 for _ in range(12))
    d2h = (ncell * 8 + 365 * len(stations) * 8) * args.members
    e2e = {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "steps": e2e_steps, "timing": "host wall clock around the API calls, max over ranks; every step copies one year of "
           "forcing from pinned host memory (for the following year, on the library's copy stream) and reads the results back"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak" if tiles == 1 else "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": ("configs[1]: 0.5deg global synthetic grid (67420 cells), daily, routing + 100-band snow; "
                                    "1 step = 1 simulated year (365 days), default 30 steps = the 30-year run") if tiles == 1 else
                                   ("configs[4] at its size: synthetic grid with the cell count of a 5-arcmin world (32 disjoint copies of the 0.5deg "
                                    "world = 2157440 cells), one member, sharded by whole drainage basin over the GPUs; 1 step = 1 simulated year"),
                       "cells": int(total_cells) if tiles > 1 else w.ng, "cells_this_rank": ncell, "members_per_gpu": args.members, "days_per_step": 365,
                       "parallelism": (f"{world} replica(s) of the single member (a 0.5deg member does not shard), no data-path collective; "
                                       "the sharded configurations are in `sharded`" if tiles == 1 else
                                       f"{world} basin shard(s) of one grid, no data-path collective"),
                       "l2": "inputs larger than L2: 183 MB state+statics per member and 394 MB of forcing per year are streamed every step (x32 for 5arcmin)",
                       "routing_levels": m.nlevels},
            "gpu_launches": int(launches), "e2e": e2e, "roofline": roofline, "clocks": clocks}
    if e2e_classes is not None:
        line["e2e_classes"] = e2e_classes
        line["config"]["inputs"] = ("synthetic world written in the reference's file formats, read and initialised by the product's host layer "
                                    "(libwghost.so: rout_prepare, init sequence of integrate_wghm_); forcing uploaded from the generator's arrays")
    m.close()
    return line


# ------------------------------------------------------------------------------------------------
# sharded legs: configs[2], configs[3], configs[4]
# ------------------------------------------------------------------------------------------------
def leg_sweep(args, D, w, ini, forcing, nsets_total=1024):
    """configs[2]: calibration sweep, parameter sets sharded over the ranks in contiguous blocks"""
    from oracle import synth_world as sw
    from watergap2_b200 import calibration as cal
    from watergap2_b200.ensemble import shard_members
    first, count = shard_members(nsets_total, D.world, D.rank)
    t_setup = time.perf_counter()
    m = make_model(w, ini, count, D.local, npset=count)
    # set i of SURVEY 8d config 3 (seeded by the GLOBAL set index): one value per set for gamma, CFA, the outflow
    # coefficients, the snow thresholds and the directly applied multipliers (device fills); the multipliers that enter
    # host-derived grids (root depth, LAI, groundwater factor, Rg_max) stay at 1
    sets = {}
    for k in range(count):
        g = first + k
        if g == 0:
            continue  # set 0 = the per-cell perturbed defaults uploaded from the host
        p = sw.default_params(w, g)
        sets[k] = {name: float(p[row][0]) for row, name in SWEEP_FIELDS.items()}
        for name, v in sets[k].items():
            m.fill(name, v, index=k)
    f = forcing[0]
    m.forcing_reserve(31)
    m.set_forcing(0, 31, f["P"], f["T"], f["SW"], f["LW"])
    stations = np.argsort(-w.acc)[:50].astype(np.int32)
    m.record_cells(stations, 31)
    m.synchronize()
    t_setup = time.perf_counter() - t_setup

    # parity sample: 3 days from the cold start, the first and the last set of this rank against the oracle
    m.step_days(1, 0, 1, 0, 3)
    pairs = []
    for k in sorted({0, count - 1}):
        ini_k = dict(ini)
        if k in sets:
            pb = np.array(ini["params"], np.float64).reshape(26, -1).copy()
            for row, name in SWEEP_FIELDS.items():
                pb[row, :] = sets[k][name]
            ini_k["params"], ini_k["gamma_hbv"], ini_k["cfa"] = pb, pb[0].copy(), pb[1].copy()
        o = oracle_for(ini_k)
        o.set_forcing_month(f)
        for d in range(1, 4):
            o.step_day(d, 0, d)
        pairs.append((f"set {first + k}", o, lambda name, k=k: m.get(name, k)))
    par = parity_sample(pairs, PARITY_NAMES)
    par["parity_sample"] = f"sets {first} and {first + count - 1} of rank {D.rank}: 3 days from the cold start at 67420 cells vs oracle/wg_oracle.c"

    gather_ms = []

    def step(k):
        m.step_days(1, 0, 1, 0, 31)
        m.synchronize()  # (the record read-back below would wait for the month anyway; done here so that gather_ms is the exchange alone)
        t0 = time.perf_counter()
        annual = np.stack([cal.annual_runoff_km3(m.get_record(31, mem), 31) for mem in range(count)])  # [sets, 1, stations]
        table = cal.gather_annual_runoff(annual, nsets_total, device="cuda")
        gather_ms.append((time.perf_counter() - t0) * 1e3)
        step.table = table

    step(0)  # warm-up (graph instantiation)
    gather_ms.clear()
    l0 = m.kernel_launches
    ms = D.timed(m, step, args.leg_steps)
    cell_days = float(w.ng) * 31 * nsets_total * args.leg_steps
    peak, _ = measured_peaks()
    per_gpu_bytes = (BYTES_VERTICAL + BYTES_ROUTING) * float(w.ng) * 31 * count * args.leg_steps
    out = {"config": "configs[2]: 0.5deg calibration sweep, 1024 parameter sets in contiguous blocks per GPU (ensemble.shard_members); set i draws "
                     "gamma, CFA, surface / groundwater outflow coefficients, snow thresholds and 5 multipliers (SURVEY 8d, seed + i); "
                     "1 step = 1 simulated month (31 days) of all sets + read-back of the 50-station record + all-gather of the annual runoff table",
           "value": cell_days / (ms / 1e3), "unit": UNIT, "scaling": "strong", "sets_total": nsets_total, "sets_this_rank": count,
           "days_per_step": 31, "steps": args.leg_steps, "ms_per_step": ms / args.leg_steps,
           "step_frac": round(per_gpu_bytes / (ms / 1e3) / 1e9 / peak, 4),
           "gather_ms_per_step": round(float(np.mean(gather_ms)), 3),
           "gathered_table_shape": list(np.asarray(step.table).shape), "collective": "all_gather_into_tensor (NCCL), inside the timed region",
           "gpu_launches": int(m.kernel_launches - l0), "setup_s": round(t_setup, 2)}
    out.update(par)
    m.close()
    return out


def leg_enkf(args, D, w, ini, forcing, nmember_total=256):
    """configs[3]: EnKF ensemble with per-member forcing; statistics of the [ncell x 10] state vector over NCCL"""
    from watergap2_b200.ensemble import ensemble_state_moments, shard_members
    first, count = shard_members(nmember_total, D.world, D.rank)
    t_setup = time.perf_counter()
    m = make_model(w, ini, count, D.local)
    base = forcing[0]
    pool = np.random.default_rng(20240607).standard_normal(base["P"].size + 4096 * nmember_total).astype(np.float32)
    pool_exp = np.exp(np.float32(0.1) * pool)

    def member_forcing(g):  # P x lognormal(sigma 0.1), T + N(0, 1 K); member g reads the pool at its own offset
        a, b = 4096 * g, 4096 * g + 1777
        n = base["P"].size
        return {"P": (base["P"].ravel() * pool_exp[a:a + n]).reshape(base["P"].shape), "T": (base["T"].ravel() + pool[b:b + n]).reshape(base["T"].shape),
                "SW": base["SW"], "LW": base["LW"]}

    m.forcing_reserve(31, per_member=True)
    for k in range(count):
        fk = member_forcing(first + k)
        m.set_forcing(0, 31, fk["P"], fk["T"], fk["SW"], fk["LW"], member=k)
        m.synchronize()  # the host arrays of this member are temporaries
    t_setup = time.perf_counter() - t_setup

    # parity sample: first and last member of this rank, 3 days from the cold start against the oracle
    m.step_days(1, 0, 1, 0, 3)
    pairs = []
    for k in sorted({0, count - 1}):
        o = oracle_for(ini)
        o.set_forcing_month(member_forcing(first + k))
        for d in range(1, 4):
            o.step_day(d, 0, d)
        pairs.append((f"member {first + k}", o, lambda name, k=k: m.get(name, k)))
    par = parity_sample(pairs, PARITY_NAMES)
    par["parity_sample"] = (f"members {first} and {first + count - 1} of rank {D.rank} with their own forcing: 3 days from the cold start at 67420 cells "
                            "vs oracle/wg_oracle.c; the moments kernel vs a sequential host sum over wgk_state_vector of this rank's members")
    # the reduction kernel against the values of wgk_state_vector, summed on the host in member order (bit-equal expected)
    m.month_begin()
    m.step_days(4, 0, 4, 3, 2)
    ps, pq, n = m.ensemble_moments("month")
    import torch
    from watergap2_b200.ensemble import device_tensor
    m.synchronize()
    got = device_tensor(ps, (2, n, 10), D.local).cpu().numpy()
    cells = np.arange(w.ng, dtype=np.int32)
    s, ss = np.zeros((n, 10)), np.zeros((n, 10))
    for k in range(count):
        v = m.state_vector(cells, "month", member=k)
        s += v
        ss += v * v
    par["moments_kernel_equals_host_sum"] = bool(np.array_equal(got[0], s) and np.array_equal(got[1], ss))

    timing = {}
    coll, stats = [], {}

    def step(k):
        m.month_begin()
        m.step_days(1, 0, 1, 0, 31)
        t = {}
        mean, var = ensemble_state_moments(m, nmember_total, "month", timing=t)
        coll.append(t.get("collective_ms", 0.0))
        timing.update(t)
        stats["mean"], stats["var"] = mean, var

    step(0)
    coll.clear()
    l0 = m.kernel_launches
    ms = D.timed(m, step, args.leg_steps)
    cell_days = float(w.ng) * 31 * nmember_total * args.leg_steps
    peak, _ = measured_peaks()
    per_gpu_bytes = (BYTES_VERTICAL + BYTES_ROUTING) * float(w.ng) * 31 * count * args.leg_steps
    out = {"config": "configs[3]: EnKF ensemble of 256 members (contiguous blocks per GPU), per-member forcing (P x lognormal 0.1, T + N(0,1 K)); "
                     "1 step = 1 assimilation cycle: wgk_month_begin, 31 simulated days, k_ensemble_moments over the monthly-mean state vector "
                     "[67420 x 10] of the rank's members, ONE in-place NCCL all-reduce of sum | sumsq (10.8 MB), k_moments_finish, mean / variance to the host",
           "value": cell_days / (ms / 1e3), "unit": UNIT, "scaling": "strong", "members_total": nmember_total, "members_this_rank": count,
           "days_per_step": 31, "steps": args.leg_steps, "ms_per_step": ms / args.leg_steps,
           "step_frac": round(per_gpu_bytes / (ms / 1e3) / 1e9 / peak, 4),
           "collective_ms_per_step": round(float(np.mean(coll)), 4), "allreduce_bytes": timing.get("allreduce_bytes"),
           "collective": "dist.all_reduce (NCCL) on the context's stream, inside the timed region" if D.world > 1 else "single rank: no collective issued",
           "ensemble_mean_snow_mm": float(stats["mean"][:, 1].mean()), "ensemble_var_soil_max": float(stats["var"][:, 2].max()),
           "gpu_launches": int(m.kernel_launches - l0), "setup_s": round(t_setup, 2)}
    out.update(par)
    m.close()
    return out


def build_basin_model(D, w, ini, forcing, tiles, members=1):
    """a grid of `tiles` disjoint copies of the 0.5 degree world; with N ranks every rank takes whole drainage basins"""
    import watergap2_b200 as wg
    from watergap2_b200.ensemble import shard_by_basin, subgrid_inputs, tile_inputs
    topo = ini["_topology"]
    fields, ro, dc = tile_inputs(ini, topo["rout_order"], topo["outflow_cell"], tiles)
    shard = None
    if D.world > 1:
        b = np.asarray(topo["basins2"]).astype(np.int64)
        basins = np.concatenate([np.where(b > 0, b + t * (int(b.max()) + 1), 0) for t in range(tiles)])
        shard = np.nonzero(shard_by_basin(basins, D.world) == D.rank)[0]
        fields, ro, dc = subgrid_inputs(fields, ro, dc, shard)
    ncell = int(np.asarray(ro).size)
    m = wg.Model(ncell, nmember=members, npset=1, device=D.local)
    m.set_topology(ro, dc, cell_class=wg.cell_classes(fields))
    m.load(fields)
    del fields

    def grid_of(a):  # a [ng][31] grid of the base world -> the rank's cells of the tiled grid
        t = np.concatenate([a] * tiles, axis=0)
        return np.ascontiguousarray(t if shard is None else t[shard])
    forcing = [{k: grid_of(v) for k, v in f.items()} for f in forcing]
    return m, ncell, forcing, shard


def leg_basins(args, D, w, ini, forcing, tiles=32):
    """configs[4]: one 2.16 M-cell grid split by whole drainage basin over the ranks (strong scaling)"""
    t_setup = time.perf_counter()
    m, ncell, f, shard = build_basin_model(D, w, ini, forcing[:1], tiles)
    f = f[0]
    m.forcing_reserve(31)
    m.set_forcing(0, 31, f["P"], f["T"], f["SW"], f["LW"])
    m.synchronize()
    t_setup = time.perf_counter() - t_setup
    cells_per_rank = D.gather(ncell)
    total_cells = float(sum(cells_per_rank))

    # parity sample: the rank's cells of ONE copy of the world, 3 days from the cold start against the oracle
    m.step_days(1, 0, 1, 0, 3)
    idx = np.arange(ncell, dtype=np.int64) if shard is None else shard
    tile_of, cell_of = idx // w.ng, idx % w.ng
    t_sample = int(np.bincount(tile_of).argmax())
    sel = np.nonzero(tile_of == t_sample)[0]
    o = oracle_for(ini)
    o.set_forcing_month(forcing[0])
    for d in range(1, 4):
        o.step_day(d, 0, d)

    class Sub:  # the oracle restricted to the sampled cells
        def field(self, name):
            a = o.field(name)
            return a.reshape(w.ng, -1)[cell_of[sel]].ravel()

    par = parity_sample([(f"copy {t_sample}", Sub(), lambda name: m.get(name).reshape(ncell, -1)[sel].ravel())], PARITY_NAMES)
    par["parity_sample"] = (f"the {sel.size} cells of copy {t_sample} of the world held by rank {D.rank}: 3 days from the cold start vs oracle/wg_oracle.c")

    m.step_days(1, 0, 1, 0, 31)  # warm-up (graph instantiation)
    l0 = m.kernel_launches
    ms = D.timed(m, lambda k: m.step_days(1, 0, 1, 0, 31), args.leg_steps)
    launches = m.kernel_launches - l0
    prof = {"vertical": [0.0, 0], "river_level": [0.0, 0], "tail": [0.0, 0], "other": [0.0, 0]}
    for d in range(5):
        p = m.profile_schedule(1 + d, 0, 1 + d, d)
        for k, (t, n) in p.items():
            prof[k][0] += t / 5
            prof[k][1] = n
    lv = np.bincount(m.levels(), minlength=m.nlevels)
    cell_days = total_cells * 31 * args.leg_steps
    peak, _ = measured_peaks()
    per_gpu_bytes = (BYTES_VERTICAL + BYTES_ROUTING) * float(ncell) * 31 * args.leg_steps
    mean_cells = total_cells / D.world
    # what does not shrink with the shard: the chain of dependency levels (one task per level and day; narrow levels in one CTA)
    narrow = int((lv <= 256).sum())
    out = {"config": "configs[4] at its size: a grid with the cell count of a 5-arcmin world (32 disjoint copies of the 0.5deg world = 2157440 cells), one member, "
                     "whole drainage basins bin-packed onto the GPUs (ensemble.shard_by_basin, LPT), no data-path collective; 1 step = 1 simulated month (31 days)",
           "value": cell_days / (ms / 1e3), "unit": UNIT, "scaling": "strong", "cells_total": int(total_cells), "cells_per_rank": [int(c) for c in cells_per_rank],
           "imbalance_max_over_mean": round(max(cells_per_rank) / mean_cells, 4), "days_per_step": 31, "steps": args.leg_steps,
           "ms_per_step": ms / args.leg_steps, "ms_per_simulated_day": ms / args.leg_steps / 31,
           "step_frac_this_rank": round(per_gpu_bytes / (ms / 1e3) / 1e9 / peak, 4),
           "routing_levels": m.nlevels, "narrow_levels": narrow, "widest_level_cells": int(lv.max()),
           "isolated_ms_per_day": {k: round(v[0], 4) for k, v in prof.items() if v[1]},
           "launches_per_day": {k: v[1] for k, v in prof.items() if v[1]},
           "limiting_kernel": "k_river_level / k_tail_chunk: the chain of %d dependency levels per simulated day does not shorten when the cells per level are "
                              "divided among the GPUs (isolated_ms_per_day.river_level + tail vs vertical); the vertical kernels scale with the cells" % m.nlevels,
           "gpu_launches": int(launches), "setup_s": round(t_setup, 2)}
    out.update(par)
    m.close()
    return out


def run_wgk(args):
    D = Dist()
    w, ini = build_inputs()
    forcing = year_forcing(w)
    line = run_headline(args, D, w, ini, forcing)
    legs = [] if args.legs in ("none", "") or args.workload != "0.5deg" or args.members != 1 else args.legs.split(",")
    if legs == ["all"]:
        legs = ["sweep", "enkf", "basins"]
    sharded = {}
    for name in legs:
        t0 = time.perf_counter()
        try:
            if name == "sweep":
                sharded["sweep_1024"] = leg_sweep(args, D, w, ini, forcing, args.sweep_sets)
            elif name == "enkf":
                sharded["enkf_256"] = leg_enkf(args, D, w, ini, forcing, args.enkf_members)
            elif name == "basins":
                sharded["basins_5arcmin"] = leg_basins(args, D, w, ini, forcing, args.basin_tiles)
            else:
                raise SystemExit(f"unknown leg {name}")
        except Exception as e:  # a failed leg must not take the headline line with it (every rank fails alike)
            if D.world > 1:
                raise
            sharded[name] = {"error": f"{type(e).__name__}: {e}"}
        key = {"sweep": "sweep_1024", "enkf": "enkf_256", "basins": "basins_5arcmin"}.get(name, name)
        if key in sharded:
            sharded[key]["leg_wall_s"] = round(time.perf_counter() - t0, 1)
    if sharded:
        line["sharded"] = sharded
    if D.rank == 0:
        line["cpu_baseline"] = cpu_baseline_sample(w, days=31) if not args.no_cpu else None
        print(json.dumps(line))
    D.close()


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline
# ------------------------------------------------------------------------------------------------
def _harness():
    p = os.path.join(ROOT, "oracle", "_ref", f"ref_harness_{NG}")
    return p if os.path.exists(p) else None


def _run_reference(w, months, extra):
    """write the world for `months` months of 1901 and run the reference replay; -> (timing dict, per-day seconds)"""
    from oracle import synth_world as sw
    tmp = tempfile.mkdtemp(prefix="wg_bench_ref_")
    sw.write_world(w, tmp, (1901, 1901), (1, months), grid_store=0, daily_discharge=False)
    dt_file = os.path.join(tmp, "day_times.txt")
    out = subprocess.run([_harness(), "replay", os.path.join(tmp, "config.txt"), "-", "--time-only", "--day-times", dt_file] + extra,
                         capture_output=True, text=True, cwd=tmp)
    mt = re.search(r"REF_TIMING (\{.*\})", out.stdout)
    if not mt:
        raise RuntimeError("reference harness failed: " + out.stdout[-400:] + out.stderr[-400:])
    times = np.loadtxt(dt_file)
    subprocess.run(["rm", "-rf", tmp])
    return json.loads(mt.group(1)), np.atleast_1d(times)


def cpu_baseline_sample(w, days=31):
    """the same day range as the first timed days of `--impl reference` (January 1901 from the cold start), so that the two
    numbers describe the same thing; the reference arm's line is THE baseline the driver compares with"""
    if _harness():
        timing, _ = _run_reference(w, 1, [])
        return {"value": timing["cell_days_per_s"], "unit": UNIT, "cores": 8, "kind": "reference",
                "sample": f"{timing['days']} simulated days (January 1901) of the same 67420-cell world through the compiled "
                          f"reference (oracle/_ref): calcNewDay on 8 OpenMP threads (hard-wired, integrateWGHM.cpp:106), "
                          f"routing serial; host has {os.cpu_count()} cores.  `bench.py --impl reference` (10-day steps over the "
                          f"months after the warm-up) is the baseline of record; this sample differs from it by the day range only",
                "t_vertical_s": timing["t_vertical_s"], "t_routing_s": timing["t_routing_s"]}
    # no compiled reference on this machine: time the C port (scalar)
    from oracle import synth_world as sw, wg_init, wgo
    ini = wg_init.derive(w)
    o = oracle_for(ini)
    o.set_forcing_month(sw.forcing_month(w, 1901, 1))
    t0 = time.perf_counter()
    for d in range(1, days + 1):
        o.step_day(d, 0, d)
    dt = time.perf_counter() - t0
    return {"value": w.ng * days / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"{days} simulated days (January 1901) of the same world through the scalar C port (oracle/wg_oracle.c)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import synth_world as sw
    w = sw.build_world(NG)
    days_per_step = 10
    total_days = (args.steps + args.warmup) * days_per_step
    months, acc = 0, 0
    while acc < total_days:
        acc += NDAYS[months]
        months += 1
    if _harness():
        timing, times = _run_reference(w, months, [])
        kind, cores = "reference", 8
    else:
        from oracle import wg_init, wgo
        ini = wg_init.derive(w)
        o = oracle_for(ini)
        times = []
        for sd in range(1, total_days + 1):
            doy, mon, dom = wgo.calendar(sd)
            if dom == 1:
                o.set_forcing_month(sw.forcing_month(w, 1901, mon + 1))
            t0 = time.perf_counter()
            o.step_day(doy, mon, dom)
            times.append(time.perf_counter() - t0)
        times = np.array(times)
        kind, cores = "port", 1
    t = times[args.warmup * days_per_step: total_days]
    secs = float(t.sum())
    value = float(w.ng) * days_per_step * args.steps / secs
    sample = (f"each step = {days_per_step} consecutive simulated days of the same 67420-cell world (day loop body of "
              f"integrateWGHM.cpp:755-798: calcNewDay on 8 OpenMP threads, routing serial, updateLandAreaFrac); "
              f"{args.warmup} warm-up + {args.steps} timed steps in one process; host has {os.cpu_count()} cores")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[1]: 0.5deg global synthetic grid (67420 cells), daily, routing + 100-band snow",
                       "cells": w.ng, "days_per_step": days_per_step},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="wgk", choices=["wgk", "reference"])
    ap.add_argument("--members", type=int, default=1, help="members (independent model runs) per GPU")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline sample")
    ap.add_argument("--workload", default="0.5deg", choices=["0.5deg", "5arcmin"],
                    help="0.5deg: BASELINE configs[1] (default, the driver's line).  5arcmin: configs[4] as the headline, a grid with the cell count "
                         "of a 5-arcmin world (32 disjoint copies of the 0.5 degree world = 2 157 440 cells), one member, sharded by whole "
                         "drainage basin over the GPUs")
    ap.add_argument("--legs", default="all", help="sharded legs measured after the headline: all | none | comma list of sweep,enkf,basins")
    ap.add_argument("--leg-steps", type=int, default=3, help="timed steps (simulated months) of each sharded leg")
    ap.add_argument("--sweep-sets", type=int, default=1024)
    ap.add_argument("--enkf-members", type=int, default=256)
    ap.add_argument("--basin-tiles", type=int, default=32)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_wgk(args)


if __name__ == "__main__":
    main()
