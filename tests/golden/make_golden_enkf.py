#!/usr/bin/env python3
"""Regenerate tests/golden/ref_ng1000_enkf.npz from the COMPILED REFERENCE (needs /root/reference).

The reference's own PDAF exchange functions are run on January 1901 of the 1000-cell world of
ref_ng1000.npz (same seed, same cold start): extract_sub_ (extractsub.cpp:17), enkf_wghmstate_
(enKF2wghmState.cpp:17) and the restore of the next cycle (routing.cpp:851 setStorages, daily.cpp:1896
setStorages), through `ref_harness replay --enkf` (oracle/ref_harness.cpp).  The "analysis" handed to
enkf_wghmstate_ is the extracted vector plus a seeded perturbation that reaches every limit of
enKF2wghmState.cpp:89-121 / 440-471 (negative storages, snow above 1000 mm, cells without snow).

Stored for the cells of the region only: the reference's extract vector, field, prediction, mean field,
its daily WghmStateFile routing entries of the month, the month-end state before the update, and the
last-day state / snow in elevation / restored in-memory state after it.  Pins oracle/enkf_bridge.py
(tests/test_enkf_bridge.py, CPU) and the CUDA bridge (GPU).
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
from oracle import synth_world as sw, wgo  # noqa: E402

NG = 1000
STATE = ["canopy", "snow", "soil", "gw", "loc_lake_stor", "loc_wetl_stor", "glo_lake_stor", "glo_wetl_stor", "res_stor", "river_stor",
         "land_area_frac", "land_area_frac_next", "land_area_frac_prev", "status_laf_next"]


def main():
    subprocess.check_call([os.path.join(ROOT, "oracle", "build_ref.sh"), str(NG)])
    tmp = tempfile.mkdtemp(prefix="wg_golden_enkf_")
    w = sw.build_world(NG)
    sw.write_world(w, tmp, (1901, 1901), (1, 1))
    cells = np.arange(5, NG, 7, dtype=np.int32)
    rng = np.random.default_rng(20240607)
    n = cells.size
    pert = rng.normal(0., 3., (n, 10)) * (rng.random((n, 10)) < 0.8) - 50. * (rng.random((n, 10)) < 0.05)
    pert[rng.random(n) < 0.1, 1] += 1500.  # snow far above the 1000 mm limit
    pert[::9] = 0.                         # cells whose analysis equals the prediction
    mean_field = rng.normal(0., 5., (n, 10))
    edir = os.path.join(tmp, "enkf")
    os.makedirs(edir)
    with open(os.path.join(edir, "ids.txt"), "w") as f:
        f.write("ID lon lat\n")
        for c in cells:
            f.write(f"{c + 1} 0.0 0.0\n")
    pert.astype("<f8").tofile(os.path.join(edir, "perturb.bin"))
    mean_field.astype("<f8").tofile(os.path.join(edir, "meanfield.bin"))
    dump = os.path.join(tmp, "dump.wgd")
    subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", f"ref_harness_{NG}"), "replay", os.path.join(tmp, "config.txt"), dump,
                           "--days", "1-31", "--snow-days", "31-31", "--enkf", edir], stdout=subprocess.DEVNULL, cwd=tmp)
    recs = wgo.read_dump(dump, days=set(range(1, 32)) | {9000})
    out = {"cells": cells, "perturb": pert, "mean_field": mean_field}
    for k in ("enkf_extract", "enkf_field", "enkf_prediction", "enkf_lastday"):
        out[k] = recs[(k, 9000)].reshape(n, 10)
    out["enkf_snow_elev"] = recs[("enkf_snow_elev", 9000)].reshape(n, 101)
    out["enkf_month_mean"] = recs[("enkf_month_mean", 9000)].reshape(NG, 10)[cells]
    out["routing_mm"] = np.stack([recs[("wghm_routing_mm", d)].reshape(7, NG)[:, cells] for d in range(1, 32)])  # [31][7][n]
    for k in STATE:
        out["before/" + k] = recs[(k, 31)][cells]
        out["after/" + k] = recs[(k, 9000)][cells]
    out["before/snow_bands"] = recs[("snow_bands", 31)].reshape(NG, 101)[cells]
    out["after/snow_bands"] = recs[("snow_bands", 9000)].reshape(NG, 101)[cells]
    path = os.path.join(ROOT, "tests", "golden", f"ref_ng{NG}_enkf.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
