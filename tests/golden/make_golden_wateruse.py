#!/usr/bin/env python3
"""Regenerate tests/golden/ref_ng1000_wateruse.npz from the COMPILED REFERENCE (needs /root/reference).

Groundwork for SURVEY 8f row 4 (water use), which is NOT on the product path yet: the 1000-cell world of ref_ng1000.npz run
for January-February 1901 with net abstractions from surface water and groundwater (subtract_use = 2, the other use options at
their canonical 0; inputs from oracle/synth_world.write_world(water_use=True)) through `ref_harness replay`, which follows
integrateWGHM.cpp:645-647 (dailyNUInit per year) and :794-796 (calcNextDay_M per day).  Stored: the monthly abstraction grids,
and the reference's state, fluxes and water-use arrays after days 1, 2, 31 and 59 - the vectors the restatement and the kernels
of that row will be pinned against.  tests/test_oracle_golden.py only checks that the fixture differs from the run without use.
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
DAYS = [1, 2, 31, 59]


def main():
    subprocess.check_call([os.path.join(ROOT, "oracle", "build_ref.sh"), str(NG)])
    tmp = tempfile.mkdtemp(prefix="wg_golden_wu_")
    w = sw.build_world(NG)
    sw.write_world(w, tmp, (1901, 1901), (1, 2), water_use=True)
    dump = os.path.join(tmp, "dump.wgd")
    subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", f"ref_harness_{NG}"), "replay", os.path.join(tmp, "config.txt"), dump,
                           "--days", "1-59"], stdout=subprocess.DEVNULL, cwd=tmp)
    recs = wgo.read_dump(dump, days=set(DAYS))
    out = {"ng": np.int32(NG), "days": np.array(DAYS, np.int32)}
    for (name, day), a in recs.items():
        if day in DAYS and a.size == NG:
            out[f"d{day}/{name}"] = a
    rd = lambda fn: np.fromfile(os.path.join(tmp, "input", fn), ">f4").astype(np.float32)
    for fn in ("G_NETUSE_SW_m3_1901.12.UNF0", "G_NETUSE_GW_m3_1901.12.UNF0", "G_IRRIG_WITHDRAWAL_USE_SW_m3_1901.12.UNF0",
               "G_IRRIG_CONS_USE_SW_m3_1901.12.UNF0", "G_FRACTRETURNGW_IRRIG.UNF0"):
        out["input/" + fn] = rd(fn)
    # unit vectors of updateNetAbstractionGW (routing.cpp:5503-5572) through `ref_harness wu_unit`: every branch
    rng = np.random.default_rng(20240607)
    n = NG
    rem = rng.normal(0., 2e-4, n) * (rng.random(n) < 0.8)
    rem[rng.random(n) < 0.05] = rng.uniform(-1e-12, 1e-12, 1)      # inside the numerical dead band
    wusi = np.where(rng.random(n) < 0.7, rng.gamma(0.8, 3e-4, n), 0.)
    cusi = wusi * rng.uniform(0.2, 0.9, n)
    frgi = rng.uniform(0.05, 0.7, n)
    uns_irr = np.where(rng.random(n) < 0.6, rng.gamma(0.8, 2e-4, n), 0.)
    uns_oth = np.where(rng.random(n) < 0.5, rng.gamma(0.8, 1e-4, n), 0.)
    red_rf = -rng.gamma(0.8, 1e-4, n) * (uns_irr > 0)
    nug = rng.normal(1e-4, 2e-4, n)
    vin = np.stack([rem, wusi, cusi, frgi, uns_irr, uns_oth, red_rf, nug])
    fin, fout = os.path.join(tmp, "wu_in.f64"), os.path.join(tmp, "wu_out.f64")
    vin.astype("<f8").tofile(fin)
    subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", f"ref_harness_{NG}"), "wu_unit", fin, fout], stdout=subprocess.DEVNULL)
    out["unit/in"] = vin
    out["unit/out"] = np.fromfile(fout, "<f8").reshape(6, n)
    path = os.path.join(ROOT, "tests", "golden", f"ref_ng{NG}_wateruse.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB", len(out), "arrays")


if __name__ == "__main__":
    main()
