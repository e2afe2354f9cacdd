#!/usr/bin/env python3
"""Short hot-path run for ncu captures: full 0.5 degree world, M members, D days, plain launches
(no graph, so that every kernel is a separate ncu launch)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--members", type=int, default=1)
    ap.add_argument("--days", type=int, default=6)
    ap.add_argument("--graph", type=int, default=0)
    ap.add_argument("--layout", default=None, choices=[None, "cells", "members"], help="force the layout of the member arrays (WGK_LAYOUT)")
    a = ap.parse_args()
    if a.layout:
        os.environ["WGK_LAYOUT"] = a.layout
        if a.layout == "members":
            os.environ["WGK_DAY_SCHEDULE"] = "wholeday"
    import watergap2_b200 as wg
    from oracle import synth_world as sw
    w, ini = bench.build_inputs()
    m = wg.Model(w.ng, nmember=a.members, use_graph=a.graph)
    topo = ini["_topology"]
    m.set_topology(topo["rout_order"], topo["outflow_cell"], cell_class=wg.cell_classes(ini))
    m.load(ini, member=0)
    for k in range(1, a.members):
        m.copy_member(0, k)
    f = sw.forcing_month(w, 1901, 1)
    m.forcing_reserve(31)
    m.set_forcing(0, 31, f["P"], f["T"], f["SW"], f["LW"])
    m.step_days(1, 0, 1, 0, a.days)
    m.synchronize()
    print("layout", m.layout, "done", m.kernel_launches, "launches; phases of one more day:", m.profile_day(a.days + 1, 0, a.days + 1, a.days))


if __name__ == "__main__":
    main()
