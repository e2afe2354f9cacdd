#!/usr/bin/env python3
"""Development tool: where a warp of the thread-per-cell vertical task spends its cycles (one member, level-0 cells), inside
the running 365-day wavefront graph and with plain launches one after the other.  Needs a library built with
-DWGK_PHASE_TIMING (tools/mkvariant.sh phase -DWGK_PHASE_TIMING; WGK_LIB=variants/libwgk_phase.so)."""
import ctypes
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
os.environ.setdefault("WGK_LIB", os.path.join(ROOT, "variants", "libwgk_phase.so"))
os.environ["WGK_VERTICAL_FORM"] = "cells"
os.environ["WGK_DAY_SCHEDULE"] = "wavefront"
import numpy as np  # noqa: E402
import bench  # noqa: E402
import watergap2_b200 as wg  # noqa: E402

w, ini = bench.build_inputs()
forcing = bench.year_forcing(w)
names = ["first loads", "tables+classify", "head (LAI, PET, canopy)", "bands", "soil loads", "soil", "(unused)"]
for use_graph in ((1, 0) if not os.environ.get("GRAPH_ONLY") else (1,)):
    topo = ini["_topology"]
    m = wg.Model(w.ng, nmember=1, use_graph=use_graph)
    m.set_topology(topo["rout_order"], topo["outflow_cell"], cell_class=wg.cell_classes(ini))
    m.load(ini)
    bench.upload_year(m, forcing)
    L = wg.lib()
    L.wgk_debug_vphases.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.wgk_debug_insitu.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    out = np.zeros((2, 2, 8), np.uint64)
    ins = (ctypes.c_ulonglong * 8)()
    m.step_days(1, 0, 1, 0, 365)
    m.synchronize()
    L.wgk_debug_vphases(m._c, out.ctypes.data)
    L.wgk_debug_insitu(m._c, ins)
    m.step_days(1, 0, 1, 0, 365)
    m.synchronize()
    L.wgk_debug_vphases(m._c, out.ctypes.data)
    L.wgk_debug_insitu(m._c, ins)
    v = list(ins)
    print(f"use_graph={use_graph}: mean duration of a vertical+local warp {v[0] / max(v[1], 1) / 1965:.1f} us, level 0 {v[4] / max(v[5], 1) / 1965:.1f} us")
    for l0 in (1, 0):
        for cls in (1, 0):
            n = float(out[l0, cls, 7]) or 1.0
            ph = out[l0, cls, :7].astype(float) / n / 1965.0
            print(f"  {'level 0' if l0 else 'levels >0'} {'band-loop warps' if cls else 'bare warps     '} n/day {n / 365:7.1f}  sum {ph.sum():6.2f} us: "
                  + ", ".join(f"{names[k]} {ph[k]:.2f}" for k in range(6)))
    wd = np.zeros((4, 1024), np.uint32)
    L.wgk_debug_warpdur.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.wgk_debug_warpdur(m._c, wd.ctypes.data)
    for k, nm in enumerate(["day 100: V part", "day 100: V + R", "day 200: V part", "day 200: V + R"]):
        x = wd[k][wd[k] > 0] / 1965.0
        if x.size:
            print(f"  level-0 warps, {nm}: n {x.size}, mean {x.mean():.1f}, p50 {np.median(x):.1f}, p90 {np.percentile(x, 90):.1f}, p99 {np.percentile(x, 99):.1f}, max {x.max():.1f} us")
    if os.environ.get("DUMP_WARPDUR"):
        order = m.device_order()           # device position of every cell
        cls = wg.cell_classes(ini)
        snowfree = m.get("s_snowfree") if m.has_field("s_snowfree") else None
        np.savez(os.path.join(ROOT, "gpurun_out", "warpdur.npz"), wd=wd, order=order, cls=cls, levels=m.levels(),
                 snow=m.get("snow"), lat=w.lat)
