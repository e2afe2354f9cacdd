#!/usr/bin/env python3
"""Development tool: how long the warps of the thread-per-cell task kernels run INSIDE the wavefront graph (all levels and
days in flight at once) against the same kernels launched one after the other (use_graph = 0).  Builds libwgk_phase.so
with -DWGK_PHASE_TIMING (tools/phase_timing.py --build-only) and reads the in-situ counters."""
import ctypes
import os
import subprocess
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
lib = os.path.join(ROOT, "watergap2_b200", "libwgk_phase.so")
if not os.path.exists(lib):
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "phase_timing.py"), "--build-only"])
os.environ["WGK_LIB"] = lib
os.environ["WGK_VERTICAL_FORM"] = "cells"
os.environ["WGK_DAY_SCHEDULE"] = "wavefront"
import bench  # noqa: E402
import watergap2_b200 as wg  # noqa: E402

w, ini = bench.build_inputs()
forcing = bench.year_forcing(w)
L = None
for use_graph in (1, 0):
    topo = ini["_topology"]
    m = wg.Model(w.ng, nmember=1, use_graph=use_graph)
    m.set_topology(topo["rout_order"], topo["outflow_cell"], cell_class=wg.cell_classes(ini))
    m.load(ini)
    bench.upload_year(m, forcing)
    L = wg.lib()
    L.wgk_debug_insitu.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    out = (ctypes.c_ulonglong * 8)()
    m.step_days(1, 0, 1, 0, 365)
    m.synchronize()
    L.wgk_debug_insitu(m._c, out)
    L.wgk_debug_stamps.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    L.wgk_debug_stamps(m._c, None, 1)
    t0 = time.perf_counter()
    m.step_days(1, 0, 1, 0, 365)
    m.synchronize()
    wall = time.perf_counter() - t0
    L.wgk_debug_insitu(m._c, out)
    v = list(out)
    us = lambda c, n: c / max(n, 1) / 1965.0
    print(f"use_graph={use_graph}: {wall * 1e3:.1f} ms per simulated year; mean warp duration: vertical+local {us(v[0], v[1]):.1f} us "
          f"(level 0: {us(v[4], v[5]):.1f}), river+post {us(v[2], v[3]):.1f} us (level 0: {us(v[6], v[7]):.1f})")
    import numpy as np
    st = np.zeros((2, 2, 512), np.uint64)
    L.wgk_debug_stamps(m._c, st.ctypes.data, 0)
    st = st[:, :, 20:360].astype(np.int64)  # steady state
    vs, ve, rs, re = st[0, 0], st[0, 1], st[1, 0], st[1, 1]
    f = lambda x: f"{np.median(x) / 1e3:.1f}"
    print(f"  level 0, us (median over days): V kernel first-start -> last-end {f(ve - vs)}, V end -> R start {f(rs - ve)}, "
          f"R kernel {f(re - rs)}, R end -> next V start {f(vs[1:] - re[:-1])}, day period {f(vs[1:] - vs[:-1])}")
    wd = np.zeros((4, 1024), np.uint32)
    L.wgk_debug_warpdur.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.wgk_debug_warpdur(m._c, wd.ctypes.data)
    nw = int((wd[0] > 0).sum())
    d = wd[:, :nw].astype(np.float64) / 1965.0
    print(f"  level-0 vertical warps ({nw}), us on days 100/101/200/300: mean {d.mean(1).round(1)}, p90 {np.percentile(d, 90, axis=1).round(1)}, max {d.max(1).round(1)}")
    top = [set(np.argsort(-d[k])[:nw // 10]) for k in range(4)]
    print(f"  slowest 10 % of the warps shared between days 100&101: {len(top[0] & top[1])}/{nw // 10}, 100&200: {len(top[0] & top[2])}/{nw // 10}, 100&300: {len(top[0] & top[3])}/{nw // 10}; "
          f"correlation of warp durations 100~101 {np.corrcoef(d[0], d[1])[0, 1]:.2f}, 100~200 {np.corrcoef(d[0], d[2])[0, 1]:.2f}")
    # what the slow warps are made of
    rank = np.asarray(m.device_order())
    cell_of_pos = np.argsort(rank)
    cls = wg.cell_classes(ini)[cell_of_pos]
    T = forcing[3]["T"][:, 9][cell_of_pos]  # day offset 100 = 10 April
    snow = m.get("snow")[cell_of_pos]
    for name, idx in (("slowest 5 %", np.argsort(-d[0])[:nw // 20]), ("fastest 50 %", np.argsort(d[0])[:nw // 2])):
        cells = np.concatenate([np.arange(i * 32, min(i * 32 + 32, 22056)) for i in idx])
        ncls = np.mean([len(set(cls[i * 32:i * 32 + 32].tolist())) for i in idx])
        print(f"  {name}: mean duration {d[0][idx].mean():.1f} us, distinct classes per warp {ncls:.1f}, share of cells with a water body {np.mean((cls[cells] & 7) > 0):.2f}, "
              f"arid {np.mean((cls[cells] & 8) > 0):.2f}, mean T {T[cells].mean():.1f}, T spread in warp {np.mean([np.ptp(T[i * 32:i * 32 + 32]) for i in idx]):.1f}, "
              f"cells with snow (end of run) {np.mean(snow[cells] > 0):.2f}")
    m.close()
