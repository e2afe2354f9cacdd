#!/usr/bin/env python3
"""Development tool: build libwgk_phase.so with -DWGK_PHASE_TIMING and print the mean SM-clock cycles per
phase of the band-parallel tile kernel over a simulated month (warm, inside the wavefront graph)."""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import watergap2_b200 as wg  # noqa: E402

lib = os.environ.get("WGK_LIB") or os.path.join(ROOT, "watergap2_b200", "libwgk_phase.so")
if not os.environ.get("WGK_LIB"):
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false",
                           "-std=c++17", "-DWGK_PHASE_TIMING", "-shared", "-Xcompiler", "-fPIC", "--cudart", "static", "-o", lib,
                           os.path.join(ROOT, "watergap2_b200", "csrc", "wgk_api.cu")])
if len(sys.argv) > 1 and sys.argv[1] == "--build-only":
    sys.exit(0)
wg.LIB_PATH = lib
from oracle import synth_world as sw, wg_init  # noqa: E402

ng = int(os.environ.get("NG", "67420"))
os.environ["WGK_VERTICAL_FORM"] = "bands"
w = sw.build_world(ng)
ini = wg_init.derive(w)
topo = ini["_topology"]
m = wg.Model(w.ng)
m.set_topology(topo["rout_order"], topo["outflow_cell"], cell_class=wg.cell_classes(ini))
m.load(ini)
f = sw.forcing_month(w, 1901, 1)
m.forcing_reserve(31)
m.set_forcing(0, 31, f["P"], f["T"], f["SW"], f["LW"])
m.step_days(1, 0, 1, 0, 31)
m.synchronize()
L = wg.lib()
out = (ctypes.c_ulonglong * 8)()
L.wgk_debug_phases(m._c, out)
m.step_days(1, 0, 1, 0, 31)
L.wgk_debug_phases(m._c, out)
v = list(out)
nb, nbare = v[6], v[7]
print(f"tiles with band loop {nb}, bare tiles {nbare}")
n = nb + nbare
print(f"mode+preload {v[0]/n:.0f} cyc/tile; head(+prefetch issue) {v[1]/n:.0f}; band slabs {v[2]/max(nb,1):.0f} (band tiles); "
      f"bare sums {v[3]/max(nbare,1):.0f} (bare tiles); tail+local {v[4]/n:.0f}   [1 us = 1965 cycles]")
