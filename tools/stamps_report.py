#!/usr/bin/env python3
"""Where a single member's simulated year goes: %globaltimer stamps of the level-0 tasks of every day inside the running
365-day graph (wgk_stamps) - durations of V(d,0) and R(d,0), the gaps between them, and the distribution of the day period."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402


def main():
    import watergap2_b200 as wg
    from oracle import synth_world as sw
    w, ini = bench.build_inputs()
    m = wg.Model(w.ng, nmember=1)
    topo = ini["_topology"]
    m.set_topology(topo["rout_order"], topo["outflow_cell"], cell_class=wg.cell_classes(ini))
    m.load(ini, member=0)
    m.forcing_reserve(365)
    slot = 0
    for mon in range(12):
        f = sw.forcing_month(w, 1901, mon + 1)
        m.set_forcing(slot, bench.NDAYS[mon], f["P"], f["T"], f["SW"], f["LW"])
        slot += bench.NDAYS[mon]
    for _ in range(2):
        m.step_days(1, 0, 1, 0, 365)
    m.synchronize()
    t0 = time.perf_counter()
    m.step_days(1, 0, 1, 0, 365)
    m.synchronize()
    plain = (time.perf_counter() - t0) * 1e3
    m.stamps(True)
    m.step_days(1, 0, 1, 0, 365)   # the graph is rebuilt with the stamp buffer in its parameters: first launch uploads it
    m.synchronize()
    m.stamps(True)                 # reset
    t0 = time.perf_counter()
    m.step_days(1, 0, 1, 0, 365)
    m.synchronize()
    stamped = (time.perf_counter() - t0) * 1e3
    st = m.stamps(False, read=True)[:, :, :365].astype(np.int64)
    vs, ve, rs, re_ = st[0, 0], st[0, 1], st[1, 0], st[1, 1]
    q = lambda x: {"mean": round(float(np.mean(x)) / 1e3, 2), "p10": round(float(np.percentile(x, 10)) / 1e3, 2),
                   "median": round(float(np.median(x)) / 1e3, 2), "p90": round(float(np.percentile(x, 90)) / 1e3, 2),
                   "max": round(float(np.max(x)) / 1e3, 2)}
    out = {"env": {k: v for k, v in os.environ.items() if k.startswith("WGK_")}, "year_ms_wall": round(plain, 3), "year_ms_wall_stamped": round(stamped, 3),
           "span_first_V_to_last_R_ms": round((re_[-1] - vs[0]) / 1e6, 3),
           "V_us": q(ve - vs), "R_us": q(re_ - rs), "gap_V_to_R_us": q(rs - ve), "gap_R_to_nextV_us": q(vs[1:] - re_[:-1]),
           "period_us": q(vs[1:] - vs[:-1]), "launches_per_year": m.kernel_launches // 4}
    print(json.dumps(out))
    np.save(os.path.join("gpurun_out", "stamps_" + (os.environ.get("STAMP_TAG") or "default") + ".npy"), st)


if __name__ == "__main__":
    main()
