#!/usr/bin/env python3
"""Regenerate tests/golden/ref_ng1000_enkf_par.npz from the COMPILED REFERENCE (needs /root/reference).

The parameter half of the reference's PDAF exchange on the world / region of make_golden_enkf.py: extract_sub_ with
calpar_size > 0 (extractsub.cpp:81-340: sub-basin means of the calibrated parameters appended to the state vector) and
enkf_wghmstate_ (enKF2wghmState.cpp:127-431: analysed parameters clamped to their range, sub-basin means for the others,
parameterJsonFile::save_cda_txt, parameterJsonFile_cda, parameterJsonFile::save), through `ref_harness replay --enkf` with the
optional calpar_* inputs (oracle/ref_harness.cpp run_enkf).

Three calibration units over the region's 143 cells (a few cells in no unit), unit 0 calibrates parameters 0, 3, 7, 25, unit 1
parameters 0 and 22, unit 2 none; the analysis perturbation pushes two of them beyond their range.

Stored: the inputs, the appended part of the extract vector, the time-evolution text file and the parameter JSON the reference
wrote (as bytes; the JSON without its creation_datetime line).  Pins wg::extract_sub_parameters / wg::enkf_parameters /
wg::parameterJsonFile of watergap2_b200/csrc/host/wg_pdaf_bridge.cpp (tests/test_host_library.py, CPU).
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


def inputs():
    cells = np.arange(5, NG, 7, dtype=np.int32)
    n = cells.size
    rng = np.random.default_rng(20240608)
    unit = rng.integers(0, 3, n)
    unit[rng.random(n) < 0.05] = -1              # cells of the region in no calibration unit
    gmi = np.zeros((3, n), np.int32)
    for u in range(3):
        gmi[u, unit == u] = cells[unit == u] + 1
    index = np.zeros((3, 26), np.int32)
    index[0, [0, 3, 7, 25]] = 1
    index[1, [0, 22]] = 1
    lo = np.full(26, -1e30)
    hi = np.full(26, 1e30)
    lo[0], hi[0] = 0.1, 5.0
    lo[3], hi[3] = 0.5, 2.0
    lo[7], hi[7] = 0.001, 0.1
    lo[22], hi[22] = 0.001, 0.1
    lo[25], hi[25] = 0.5, 2.0
    rng_ = np.stack([lo, hi])                     # [2][26]
    pert = np.array([0.37, 5.0, -0.5, 0.0123456789, -10.0, 0.0031415926])  # 3 (root depth) above, 7 below, gamma of unit 1 below
    return cells, gmi, index, rng_, pert


def main():
    subprocess.check_call([os.path.join(ROOT, "oracle", "build_ref.sh"), str(NG)])
    tmp = tempfile.mkdtemp(prefix="wg_golden_enkfpar_")
    w = sw.build_world(NG)
    sw.write_world(w, tmp, (1901, 1901), (1, 1))
    cells, gmi, index, rng_, pert = inputs()
    n = cells.size
    edir = os.path.join(tmp, "enkf")
    os.makedirs(edir)
    with open(os.path.join(edir, "ids.txt"), "w") as f:
        f.write("ID lon lat\n")
        for c in cells:
            f.write(f"{c + 1} 0.0 0.0\n")
    np.zeros((n, 10)).astype("<f8").tofile(os.path.join(edir, "perturb.bin"))
    np.zeros((n, 10)).astype("<f8").tofile(os.path.join(edir, "meanfield.bin"))
    index.astype("<i4").tofile(os.path.join(edir, "calpar_index.bin"))
    gmi.astype("<i4").tofile(os.path.join(edir, "groupmatrixindex.bin"))
    rng_.astype("<f8").tofile(os.path.join(edir, "calpar_range.bin"))
    pert.astype("<f8").tofile(os.path.join(edir, "calpar_perturb.bin"))
    with open(os.path.join(edir, "arcid_gcrc.txt"), "w") as f:   # parameterJsonFile::save: header line, then "arcid gcrc"
        f.write("arcid gcrc\n")
        for c in range(NG):
            f.write(f"{100000 + 3 * c} {c + 1}\n")
    dump = os.path.join(tmp, "dump.wgd")
    subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", f"ref_harness_{NG}"), "replay", os.path.join(tmp, "config.txt"), dump,
                           "--days", "31-31", "--enkf", edir], stdout=subprocess.DEVNULL, cwd=tmp)
    recs = wgo.read_dump(dump, days={9000})
    out = {"cells": cells, "groupmatrixindex": gmi, "calpar_index": index, "calpar_range": rng_, "calpar_perturb": pert,
           "par_extract": recs[("enkf_par_extract", 9000)], "par_field": recs[("enkf_par_field", 9000)],
           "parameters_in": np.fromfile(os.path.join(tmp, "parameters.f64"), "<f8").reshape(26, NG)}
    txt = [f for f in os.listdir(edir) if f.startswith("calpar_") and f.endswith(".txt")]
    assert len(txt) == 1, txt
    out["cda_txt_name"] = np.frombuffer(txt[0].encode(), np.uint8)
    out["cda_txt"] = np.frombuffer(open(os.path.join(edir, txt[0]), "rb").read(), np.uint8)
    js = open(os.path.join(edir, "parameters_out.json"), "rb").read().split(b"\n")
    js = [l for l in js if not l.startswith(b'"creation_datetime"')]
    out["json"] = np.frombuffer(b"\n".join(js), np.uint8)
    path = os.path.join(ROOT, "tests", "golden", f"ref_ng{NG}_enkf_par.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e3, "kB")


if __name__ == "__main__":
    main()
