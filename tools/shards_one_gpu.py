#!/usr/bin/env python3
"""Development tool: the single 0.5 degree member as K basin shards (whole drainage basins, bit-identical to the full grid) that
step CONCURRENTLY on one GPU, each in its own context / stream: how the simulated year scales with K (premise check)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402


def main():
    import watergap2_b200 as wg
    from oracle import synth_world as sw
    from watergap2_b200.ensemble import shard_by_basin, subgrid_inputs
    w, ini = bench.build_inputs()
    topo = ini["_topology"]
    fields = {k: v for k, v in ini.items() if not k.startswith("_")}
    basins = np.asarray(topo["basins2"]).astype(np.int64)
    forcing = [sw.forcing_month(w, 1901, mon + 1) for mon in range(12)]
    for K in [int(x) for x in (os.environ.get("KS") or "1,2,4,8").split(",")]:
        part = shard_by_basin(basins, K)
        models = []
        for k in range(K):
            cells = np.nonzero(part == k)[0]
            f, ro, dc = subgrid_inputs(ini, topo["rout_order"], topo["outflow_cell"], cells) if K > 1 else (ini, topo["rout_order"], topo["outflow_cell"])
            m = wg.Model(int(np.asarray(ro).size), nmember=1)
            m.set_topology(ro, dc, cell_class=wg.cell_classes(f))
            m.load(f)
            m.forcing_reserve(365)
            slot = 0
            for mon in range(12):
                g = forcing[mon]
                sel = (lambda a: np.ascontiguousarray(a[cells])) if K > 1 else (lambda a: a)
                m.set_forcing(slot, bench.NDAYS[mon], sel(g["P"]), sel(g["T"]), sel(g["SW"]), sel(g["LW"]))
                slot += bench.NDAYS[mon]
            models.append(m)
        for _ in range(3):
            for m in models:
                m.step_days(1, 0, 1, 0, 365)
        for m in models:
            m.synchronize()
        t0 = time.perf_counter()
        n = 5
        for _ in range(n):
            for m in models:
                m.step_days(1, 0, 1, 0, 365)
        for m in models:
            m.synchronize()
        dt = (time.perf_counter() - t0) / n
        print(f"K {K}: cells {[m.ncell for m in models]} levels {[m.nlevels for m in models]}  {dt * 1e3:.2f} ms per simulated year  {w.ng * 365 / dt:.3e} cell-days/s", flush=True)
        for m in models:
            m.close()


if __name__ == "__main__":
    main()
