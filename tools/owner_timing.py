#!/usr/bin/env python3
"""Where the warps of the cell-owner schedule (k_days_owner) spend their time: per-warp cycle counts of
{vertical + local routing, hand-off waits, river + release, post-pass}, summed over the days of one call.

  WGK_OWNER_TIMING=1 python tools/owner_timing.py --days 124
"""
import argparse
import ctypes
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
os.environ.setdefault("WGK_OWNER_TIMING", "1")
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--days", type=int, default=124)
    ap.add_argument("--members", type=int, default=1)
    a = ap.parse_args()
    import watergap2_b200 as wg
    from oracle import synth_world as sw
    w, ini = bench.build_inputs()
    m = wg.Model(w.ng, nmember=a.members, use_graph=1)
    topo = ini["_topology"]
    m.set_topology(topo["rout_order"], topo["outflow_cell"], cell_class=wg.cell_classes(ini))
    m.load(ini)
    f = sw.forcing_month(w, 1901, 1)
    m.forcing_reserve(31)
    m.set_forcing(0, 31, f["P"], f["T"], f["SW"], f["LW"])
    m.step_days(1, 0, 1, 0, 31)  # warm-up
    m.synchronize()
    import time
    t0 = time.perf_counter()
    m.step_days(1, 0, 1, 0, a.days)
    m.synchronize()
    wall = time.perf_counter() - t0
    L = wg.lib()
    L.wgk_debug_owner_timing.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    nmax = 1 << 16
    t = np.zeros((nmax, 4), np.int64)
    wb = np.zeros(nmax, np.int32)
    n = L.wgk_debug_owner_timing(m._c, t.ctypes.data, wb.ctypes.data, nmax)
    if n <= 0:
        print("no timing (WGK_OWNER_TIMING unset or the owner schedule was not used)")
        return
    t, wb = t[:n] / a.days / 1965.0, wb[:n]  # us per day at 1965 MHz
    levels = np.asarray(m.levels())[np.argsort(m.device_order())] if hasattr(m, "levels") else None
    print(f"{a.days} days, {n} warps, wall {wall * 1e3:.2f} ms = {wall * 1e6 / a.days:.1f} us/day")
    names = ["vertical+local", "wait", "river+release", "post"]
    print("mean us/day per warp:", {k: round(float(v), 2) for k, v in zip(names, t.mean(0))}, "total", round(float(t.sum(1).mean()), 2))
    work = t[:, 0] + t[:, 2] + t[:, 3]
    print("work us/day percentiles 50/90/99/max:", np.percentile(work, [50, 90, 99, 100]).round(1))
    for k, nm in enumerate(names):
        print(f"  {nm:16s} p50 {np.percentile(t[:, k], 50):7.1f}  p90 {np.percentile(t[:, k], 90):7.1f}  p99 {np.percentile(t[:, k], 99):7.1f}  max {t[:, k].max():7.1f}")
    order = np.argsort(-work)[:12]
    print("slowest warps (warp, first cell, work, vertical, wait, river, post):")
    for i in order:
        print(f"  {i:5d} {wb[i]:6d} {work[i]:7.1f} {t[i, 0]:7.1f} {t[i, 1]:7.1f} {t[i, 2]:7.1f} {t[i, 3]:7.1f}")
    # per-warp vertical time against the classes of its cells (bit 0 local lake, 1 local wetland, 2 global body, 3 arid)
    rank = np.asarray(m.device_order())
    cell_of_pos = np.argsort(rank)
    cls = wg.cell_classes(ini)[cell_of_pos]
    lvl = np.asarray(m.levels())[cell_of_pos]
    we = np.append(wb[1:], w.ng)
    sig = {}
    for i in range(n):
        c = cls[wb[i]:min(we[i], wb[i] + 32)]
        key = (len(set(c.tolist())), int(np.bitwise_or.reduce(c)) & 7, len(set((c >> 3).tolist())))
        sig.setdefault(key, []).append(t[i, 0])
    print("vertical+local us/day by (distinct classes, union of water-body bits, arid variants): count mean max")
    for key in sorted(sig):
        v = np.array(sig[key])
        print(f"  {key}: {v.size:5d} {v.mean():7.1f} {v.max():7.1f}")
    print("slowest 12 by vertical: warp level ncells classes V")
    for i in np.argsort(-t[:, 0])[:12]:
        c = cls[wb[i]:min(we[i], wb[i] + 32)]
        print(f"  {i:5d} L{lvl[wb[i]]:2d} n={c.size:2d} {sorted(set(c.tolist()))} {t[i, 0]:6.1f}")
    # by position in the warp list (= by level): 16 groups
    g = np.array_split(np.arange(n), 16)
    print("by warp index group (levels ascend): mean vertical / wait / river / post")
    for idx in g:
        print(f"  warps {idx[0]:5d}-{idx[-1]:5d}: " + " ".join(f"{t[idx, k].mean():7.1f}" for k in range(4)))


if __name__ == "__main__":
    main()
