#!/usr/bin/env python3
"""bench.py — simulated cell-days per second of the WaterGAP2 daily hot path on B200.

  python bench.py --gpus N --steps K --warmup W [--impl wgk|reference] [--members M]

Workload (BASELINE.json configs[1]): the 0.5 degree global synthetic grid (67 420 cells, seed
20240607), daily time step with routing and 100 elevation-band snow.  One STEP is one
simulated model year (365 days); the default K = 30 timed steps is the 30-year run of the
config.  Per GPU one member (a single model run) unless --members is given; at N > 1 every
rank runs its own member(s) ("ensemble throughput", weak scaling, no data-path collective).

value   = cells x 365 x members x N x K / max-over-ranks(device time of the K steps), state and a
          full year of forcing resident in HBM (forcing slots are cycled on the device).
e2e     = the same metric through the C ABI with HOST buffers: every step the year's forcing
          (12 x 4 grids [ncell][31] float32, the reference's .31 layout) is copied from pinned
          host memory and packed on the device, the year is stepped, and the daily discharge of
          50 station cells plus the last day's per-cell discharge are read back to the host.
roofline, cpu_baseline: see DESIGN.md §5.

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
METRIC = "simulated cell-days/sec, 0.5deg global grid"
UNIT = "cell-days/s"


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


def make_model(w, ini, members, device):
    import watergap2_b200 as wg
    m = wg.Model(w.ng, nmember=members, npset=1, device=device)
    topo = ini["_topology"]
    m.set_topology(topo["rout_order"], topo["outflow_cell"], cell_class=wg.cell_classes(ini))
    m.load(ini)
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


def run_wgk(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the wgk path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w, ini = build_inputs()
    forcing = year_forcing(w)
    ncell, tiles, shard = w.ng, 1, None
    if args.workload == "5arcmin":
        # 32 copies of the 0.5 degree world as one grid; with N GPUs every rank takes whole drainage basins
        import watergap2_b200 as wg
        from watergap2_b200.ensemble import shard_by_basin, subgrid_inputs, tile_inputs
        tiles = 32
        topo = ini["_topology"]
        fields, ro, dc = tile_inputs(ini, topo["rout_order"], topo["outflow_cell"], tiles)
        if world > 1:
            b = np.asarray(topo["basins2"]).astype(np.int64)
            basins = np.concatenate([np.where(b > 0, b + t * (int(b.max()) + 1), 0) for t in range(tiles)])
            shard = np.nonzero(shard_by_basin(basins, world) == rank)[0]
            fields, ro, dc = subgrid_inputs(fields, ro, dc, shard)
        ncell = int(np.asarray(ro).size)
        m = wg.Model(ncell, nmember=args.members, npset=1, device=local)
        m.set_topology(ro, dc, cell_class=wg.cell_classes(fields))
        m.load(fields)
        del fields

        def grid_of(a):  # a [ng][31] grid of the base world -> the rank's cells of the tiled grid
            t = np.concatenate([a] * tiles, axis=0)
            return np.ascontiguousarray(t if shard is None else t[shard])
        forcing = [{k: grid_of(v) for k, v in f.items()} for f in forcing]
    else:
        m = make_model(w, ini, args.members, local)
    upload_year(m, forcing)
    stream = torch.cuda.ExternalStream(m.stream, device=local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ------------------------------------------------
    for _ in range(args.warmup):
        m.step_days(1, 0, 1, 0, 365)
    m.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = m.kernel_launches
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        m.step_days(1, 0, 1, 0, 365)
    e1.record(stream)
    m.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = m.kernel_launches - l0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    ncell_all = torch.tensor([float(ncell)], device="cuda", dtype=torch.float64)  # ranks hold different shards
    if world > 1 and args.workload == "5arcmin":
        dist.all_reduce(ncell_all, op=dist.ReduceOp.SUM)
    elif world > 1:
        ncell_all *= world
    total_cells = float(ncell_all.item())
    cell_days = total_cells * 365 * args.members * args.steps
    value = cell_days / (ms_max / 1e3)

    # ---- per-kernel roofline (CUDA events between the phases, plain launches) ------------------
    prof = {"vertical": 0.0, "route_local": 0.0, "route_levels": 0.0, "route_tail": 0.0, "route_post": 0.0, "day": 0.0}
    nprof = 20
    for d in range(nprof):
        p = m.profile_day(1 + d, 0, 1 + d, d)
        for k in prof:
            prof[k] += p[k] / nprof
    peak, peak_src = measured_peaks()
    bytes_v = BYTES_VERTICAL * ncell * args.members
    ach_v = bytes_v / (prof["vertical"] * 1e-3) / 1e9
    t_rout = prof["route_local"] + prof["route_levels"] + prof["route_tail"] + prof["route_post"]
    ach_r = BYTES_ROUTING * ncell * args.members / (t_rout * 1e-3) / 1e9
    form = os.environ.get("WGK_VERTICAL_FORM") or ("bands" if ncell * args.members < 32768 else "cells")
    kname = {"cells": "k_vertical_tpc", "bands": "k_vertical<VCfgSmall>", "bands2": "k_vertical<VCfgMid>"}[form]
    dominant = kname if prof["vertical"] >= t_rout else "routing sweep (k_route_local + k_route_level x L + k_route_tail)"
    # DRAM traffic of the same kernel from the committed ncu --set full capture (profiles/traffic.json, written by
    # tools/ncu_summary.py traffic ...): bytes per launch at this member count, or null when there is no capture
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            traffic = json.load(fh).get(kname.split("<")[0], {}).get(str(args.members)) if tiles == 1 else None
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": kname + " (vertical balance of the whole grid; the same device code runs per routing level "
                                            "inside k_cells_pre* in the timed graph)",
                "achieved": round(ach_v, 1), "peak": peak, "unit": "GB/s",
                "frac": round(ach_v / peak, 4), "traffic": traffic,
                "dram_achieved": round(traffic / (prof["vertical"] * 1e-3) / 1e9, 1) if traffic else None,
                "dram_frac": round(traffic / (prof["vertical"] * 1e-3) / 1e9 / peak, 4) if traffic else None,
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_v, "avg_launch_ms": round(prof["vertical"], 5),
                "note": "algorithmic bytes count every cell's 100 snow bands (SURVEY 8d: 2099 B per cell-day); cells without snow and above "
                        "freezing skip the band loop, so the measured DRAM traffic is lower than the algorithmic bytes",
                "share_of_day": round(prof["vertical"] / prof["day"], 4),
                # the whole timed step against the same peak: all kernels of a simulated year, (2099 + 617) B per cell-day
                # (per GPU)
                "step_achieved": round((BYTES_VERTICAL + BYTES_ROUTING) * cell_days / world / (ms_max / 1e3) / 1e9, 1),
                "step_frac": round((BYTES_VERTICAL + BYTES_ROUTING) * cell_days / world / (ms_max / 1e3) / 1e9 / peak, 4),
                "dominant_by_time": dominant,
                "routing": {"achieved": round(ach_r, 1), "frac": round(ach_r / peak, 4), "ms_per_day": round(t_rout, 5),
                            "levels": m.nlevels, "note": "latency-bound dependency chain at 1 member"},
                "phase_ms_per_day": {k: round(v, 5) for k, v in prof.items()}}

    # ---- end to end through the C ABI with host buffers ----------------------------------------
    # One step = one simulated year as a user of the API runs it: the year's forcing (12 months x 4
    # grids in the reference's [ncell][31] float layout) is copied from pinned host memory and packed
    # on the device, the 365 days are stepped, and the daily discharge at 50 station cells plus the
    # per-cell discharge field of the last day are read back to the host.
    pinned = []
    for mon in range(12):
        d = {}
        for k in ("P", "T", "SW", "LW"):
            d[k] = torch.from_numpy(np.ascontiguousarray(forcing[mon][k])).pin_memory()
        pinned.append(d)
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
    barrier()
    t0 = time.perf_counter()
    for y in range(1, e2e_steps + 1):
        e2e_year(y)
    m.synchronize()
    barrier()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_val = total_cells * 365 * args.members * e2e_steps / float(tt.item())
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
                       "parallelism": (f"{world} independent member shard(s), no data-path collective" if tiles == 1 else
                                       f"{world} basin shard(s) of one grid, no data-path collective"),
                       "l2": "inputs larger than L2: 183 MB state+statics per member and 394 MB of forcing per year are streamed every step (x32 for 5arcmin)",
                       "routing_levels": m.nlevels},
            "gpu_launches": int(launches), "e2e": e2e, "roofline": roofline, "clocks": clocks}
    if rank == 0:
        line["cpu_baseline"] = cpu_baseline_sample(w, days=31) if not args.no_cpu else None
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


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
    if _harness():
        timing, _ = _run_reference(w, 1, [])
        return {"value": timing["cell_days_per_s"], "unit": UNIT, "cores": 8, "kind": "reference",
                "sample": f"{timing['days']} simulated days (January 1901) of the same 67420-cell world through the compiled "
                          f"reference (oracle/_ref): calcNewDay on 8 OpenMP threads (hard-wired, integrateWGHM.cpp:106), "
                          f"routing serial; host has {os.cpu_count()} cores",
                "t_vertical_s": timing["t_vertical_s"], "t_routing_s": timing["t_routing_s"]}
    # no compiled reference on this machine: time the C port (scalar)
    from oracle import synth_world as sw, wg_init, wgo
    ini = wg_init.derive(w)
    o = wgo.Oracle(w.ng)
    for k, v in ini.items():
        if not k.startswith("_") and o.has(k):
            o.set(k, v)
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
        o = wgo.Oracle(w.ng)
        for k, v in ini.items():
            if not k.startswith("_") and o.has(k):
                o.set(k, v)
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
                    help="0.5deg: BASELINE configs[1] (default, the driver's line).  5arcmin: configs[4], a grid with the cell count of a "
                         "5-arcmin world (32 disjoint copies of the 0.5 degree world = 2 157 440 cells), one member, sharded by whole "
                         "drainage basin over the GPUs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_wgk(args)


if __name__ == "__main__":
    main()
