#!/usr/bin/env python3
"""Development tool: simulated year of a small world (default 3000 cells), both forms of the vertical kernel."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ng", type=int, default=3000)
    ap.add_argument("--members", type=int, default=1)
    a = ap.parse_args()
    from oracle import synth_world as sw, wg_init
    import watergap2_b200 as wg
    w = sw.build_world(a.ng)
    ini = wg_init.derive(w)
    topo = ini["_topology"]
    f = sw.forcing_month(w, 1901, 1)
    for form in ("bands", "bands2", "cells"):
        os.environ["WGK_VERTICAL_FORM"] = form
        m = wg.Model(w.ng, nmember=a.members)
        m.set_topology(topo["rout_order"], topo["outflow_cell"], cell_class=wg.cell_classes(ini))
        m.load(ini)
        m.forcing_reserve(31)
        m.set_forcing(0, 31, f["P"], f["T"], f["SW"], f["LW"])
        for _ in range(2):
            m.step_days(1, 0, 1, 0, 365)
        m.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            m.step_days(1, 0, 1, 0, 365)
        m.synchronize()
        dt = (time.perf_counter() - t0) / 3
        print(f"ng {a.ng} members {a.members} form {form}: {dt*1e3:.2f} ms/yr  {w.ng*365*a.members/dt:.3e} cell-days/s  levels {m.nlevels}")


if __name__ == "__main__":
    main()
