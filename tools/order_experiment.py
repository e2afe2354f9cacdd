#!/usr/bin/env python3
"""Development tool: the device order inside a dependency level is free (wgk_set_cell_classes); time a simulated year for
several sort keys - water-body class (the default), and class combined with bins of the annual mean temperature."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
import watergap2_b200 as wg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--members", type=int, default=1)
    a = ap.parse_args()
    w, ini = bench.build_inputs()
    forcing = bench.year_forcing(w)
    tmean = np.mean([f["T"][:, :28].mean(1) for f in forcing], axis=0)
    tmin = np.min([f["T"][:, :28].mean(1) for f in forcing], axis=0)
    cls = wg.cell_classes(ini).astype(np.int64)

    def bins(x, n):
        q = np.quantile(x, np.linspace(0, 1, n + 1)[1:-1])
        return np.searchsorted(q, x)

    keys = {
        "class (default)": cls,
        "none": np.zeros_like(cls),
        "class*16 + tmean16": cls * 16 + bins(tmean, 16),
        "tmean16*16 + class": bins(tmean, 16) * 16 + cls,
        "tmean8*16 + class": bins(tmean, 8) * 16 + cls,
        "tmean256": bins(tmean, 256),
        "tmin256": bins(tmin, 256),
        "tmin16*16 + class": bins(tmin, 16) * 16 + cls,
    }
    topo = ini["_topology"]
    for name, key in keys.items():
        m = wg.Model(w.ng, nmember=a.members)
        m.set_topology(topo["rout_order"], topo["outflow_cell"], cell_class=key.astype(np.uint8))
        m.load(ini)
        bench.upload_year(m, forcing)
        for _ in range(3):
            m.step_days(1, 0, 1, 0, 365)
        m.synchronize()
        t0 = time.perf_counter()
        n = 5 if a.members == 1 else 1
        for _ in range(n):
            m.step_days(1, 0, 1, 0, 365)
        m.synchronize()
        ms = (time.perf_counter() - t0) / n * 1e3
        print(f"members {a.members} key {name:22s}: {ms:8.2f} ms per simulated year = {ms / 365 / a.members * 1e3:.1f} us per member-day", flush=True)
        m.close()


if __name__ == "__main__":
    main()
