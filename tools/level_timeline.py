#!/usr/bin/env python3
"""Development tool: timeline of the fused (day, level) tasks of a single member inside the running 365-day graph, level by
level (WGK_STAMP_LEVEL): task duration, gap to the same level's next day, period, and when each level starts / finishes the year."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402


def main():
    import watergap2_b200 as wg
    from oracle import synth_world as sw
    w, ini = bench.build_inputs()
    topo = ini["_topology"]
    levels = [int(x) for x in (os.environ.get("LEVELS") or "0,1,3,8,13,20,30,45,56").split(",")]
    rows = []
    for lv in levels:
        os.environ["WGK_STAMP_LEVEL"] = str(lv)
        m = wg.Model(w.ng, nmember=1)
        m.set_topology(topo["rout_order"], topo["outflow_cell"], cell_class=wg.cell_classes(ini))
        if lv >= m.nlevels:
            m.close()
            continue
        ncl = int(np.bincount(m.levels(), minlength=m.nlevels)[lv])
        m.load(ini, member=0)
        m.forcing_reserve(365)
        slot = 0
        for mon in range(12):
            f = sw.forcing_month(w, 1901, mon + 1)
            m.set_forcing(slot, bench.NDAYS[mon], f["P"], f["T"], f["SW"], f["LW"])
            slot += bench.NDAYS[mon]
        m.stamps(True)
        for _ in range(2):
            m.step_days(1, 0, 1, 0, 365)
        m.synchronize()
        m.stamps(True)
        m.step_days(1, 0, 1, 0, 365)
        m.synchronize()
        st = m.stamps(False, read=True)[:, :, :365].astype(np.int64)
        vs, re_ = st[0, 0], st[1, 1]
        t0 = vs[0] if lv == levels[0] else None
        dur, gap, per = (re_ - vs) / 1e3, (vs[1:] - re_[:-1]) / 1e3, (vs[1:] - vs[:-1]) / 1e3
        rows.append({"level": lv, "cells": ncl, "task_us": [round(float(np.percentile(dur, q)), 1) for q in (50, 90, 99)],
                     "gap_us": [round(float(np.percentile(gap, q)), 1) for q in (50, 90, 99)],
                     "period_us": [round(float(np.percentile(per, q)), 1) for q in (50, 90, 99)], "period_mean_us": round(float(per.mean()), 2),
                     "year_span_ms": round(float(re_[-1] - vs[0]) / 1e6, 3)})
        print(json.dumps(rows[-1]), flush=True)
        m.close()


if __name__ == "__main__":
    main()
