#!/usr/bin/env python3
"""Regenerate tests/golden/ref_ng1000.npz from the COMPILED REFERENCE (needs /root/reference).

  1. builds oracle/_ref/ref_harness_1000 (reference sources, ng overridden to 1000),
  2. writes the synthetic mini world (oracle/synth_world.py, ng=1000, seed 20240607),
  3. runs the reference driver replay for Jan-Feb 1901 and dumps in-memory doubles,
  4. stores: every day-0 record (derived statics + initial state as the reference computed
     them), the forcing of both months, and the reference's state/fluxes after selected days.

The fixture pins the oracle (tests/test_oracle_golden.py) and, through the oracle-free
comparison in tests/test_gpu_parity.py, the CUDA path, on machines without the reference.
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
DAYS = [1, 2, 15, 31, 32, 45, 59]
SNOW_DAY = 59


def main():
    subprocess.check_call([os.path.join(ROOT, "oracle", "build_ref.sh"), str(NG)])
    tmp = tempfile.mkdtemp(prefix="wg_golden_")
    w = sw.build_world(NG)
    sw.write_world(w, tmp, (1901, 1901), (1, 2))
    dump = os.path.join(tmp, "dump.wgd")
    subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", f"ref_harness_{NG}"), "replay",
                           os.path.join(tmp, "config.txt"), dump, "--days", "1-59", "--snow-days", f"{SNOW_DAY}-{SNOW_DAY}"],
                          stdout=subprocess.DEVNULL, cwd=tmp)
    recs = wgo.read_dump(dump, days=set([0] + DAYS))
    out = {"ng": np.int32(NG), "days": np.array(DAYS, np.int32)}
    for (name, day), a in recs.items():
        out[f"d{day}/{name}"] = a
    for mth in (1, 2):
        f = sw.forcing_month(w, 1901, mth)
        for k, v in f.items():
            out[f"forcing{mth}/{k}"] = v
    topo = wgo.rout_prepare(w.flowdir, w.row, w.col, w.gcrc.T)
    # the reference's own routing files, as golden integers for the topology restatement
    rd = lambda name, dt: np.fromfile(os.path.join(tmp, "routing", name), dtype=np.dtype(dt).newbyteorder(">")).astype(dt)
    for fname, dt in [("G_LDD_2.UNF1", "i1"), ("G_INFLC.9.UNF4", "i4"), ("G_FLOW_ACC.UNF2", "i2"),
                      ("G_CELLS_TO_OUTLET.UNF2", "u2"), ("G_BASINS.UNF2", "u2"), ("G_BASINS_2.UNF2", "u2"),
                      ("G_OUTFLC.UNF4", "i4"), ("G_ROUT_ORDER.UNF4", "i4"), ("G_RIVERSLOPE.UNF0", "f4"),
                      ("G_RIVER_LENGTH.UNF0", "f4"), ("G_ALLOC_COEFF.5.UNF0", "f4"), ("G_START_MONTH.UNF1", "i1")]:
        out["routing/" + fname] = rd(fname, dt)
    path = os.path.join(ROOT, "tests", "golden", f"ref_ng{NG}.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")

    # second fixture: the same world started with 0..1400 mm of snow in the elevation bands of every
    # third cell (harness option --deep-snow), so that the 1000 mm cap of daily.cpp:958-976 and the
    # melt/sublimation branches of deep snow packs are pinned; statics and forcing are those above
    dump2 = os.path.join(tmp, "dump_deep.wgd")
    subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", f"ref_harness_{NG}"), "replay",
                           os.path.join(tmp, "config.txt"), dump2, "--days", "1-6", "--snow-days", "1-6", "--deep-snow"],
                          stdout=subprocess.DEVNULL, cwd=tmp)
    recs = wgo.read_dump(dump2, days={0, 1, 2, 6})
    deep = {"days": np.array([1, 2, 6], np.int32)}
    for (name, day), a in recs.items():
        if day == 0 and name not in ("snow_bands", "snow"):
            continue  # as in ref_ng1000.npz
        deep[f"d{day}/{name}"] = a
    path = os.path.join(ROOT, "tests", "golden", f"ref_ng{NG}_deepsnow.npz")
    np.savez_compressed(path, **deep)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
