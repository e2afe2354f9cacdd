#!/usr/bin/env python3
"""Regenerate tests/golden/ref_calibration.json from the COMPILED REFERENCE (needs /root/reference).

`ref_harness calib <dir>` (oracle/ref_harness.cpp) runs the calibration loop of integrate_wghm_ around the reference's own
calibGammaClass; the model run is replaced by a closed-form response of the annual station discharge to gamma,
    simulated[i] = (float)(base[i] * (s0 + s1 / (1 + gamma))),
so that every branch of findNewGamma / writeCorrFactors / createCorrectionGrid can be reached: bisection to the 1 %
criterion, gamma at the upper / lower limit with the 10 % criterion met or not (CFA, correction grid, CFS), years
without observation.  Stored: the inputs, the return value and public state after every findNewGamma call, the data
lines of CALIBRATION.OUT and STAT_CORR_FACTOR.OUT, CALIBSTATUS.OUT and the correction grid the reference wrote.
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
NG = 1000

SCENARIOS = {
    "bisection": dict(gamma0=1.0, s0=0.6, s1=1.32, skip_years=[1993]),
    "first_call_ok": dict(gamma0=2.0, s0=1.0, s1=0.0, skip_years=[]),
    "upper_limit_10pct_ok": dict(gamma0=2.0, s0=0.95, s1=0.6, skip_years=[]),
    "upper_limit_cfa": dict(gamma0=2.0, s0=1.2, s1=0.6, skip_years=[1991, 1996]),
    "lower_limit_cfa": dict(gamma0=1.0, s0=0.5, s1=0.3, skip_years=[]),
    "lower_limit_10pct_ok": dict(gamma0=1.0, s0=0.68, s1=0.3, skip_years=[]),
    "start_at_upper_limit": dict(gamma0=5.0, s0=1.2, s1=0.6, skip_years=[]),
}


def main():
    subprocess.check_call([os.path.join(ROOT, "oracle", "build_ref.sh"), str(NG)])
    exe = os.path.join(ROOT, "oracle", "_ref", f"ref_harness_{NG}")
    rng = np.random.default_rng(20240607)
    y0, y1, station = 1990, 1997, 7
    ny = y1 - y0 + 1
    obs_m3s = np.round(rng.uniform(800., 1500., ny), 2)           # RIVER.DAT, m3/s
    base = obs_m3s * 365 * 24 * 60 * 60 / 1e9 * rng.uniform(0.93, 1.07, ny)  # km3/year around the observation
    inflow = rng.uniform(2., 5., ny)
    use = rng.uniform(0.1, 1.0, ny)
    sbasin = np.where(rng.random(NG) < 0.2, station, 3).astype("<i2")
    pot = rng.normal(0.02, 0.05, (ny, NG)).astype(np.float32)      # G_POT_CELL_RUNOFF_<year>.UNF0
    pot[:, ::17] = 0.
    out = {"eval_start_year": y0, "end_year": y1, "station": station, "observed_m3s": obs_m3s.tolist(), "base": base.tolist(),
           "inflow": inflow.tolist(), "water_use": use.tolist(), "sbasin": sbasin.tolist(),
           "pot_cell_runoff": [[float(v) for v in row] for row in pot], "scenarios": {}}
    for name, sc in SCENARIOS.items():
        d = tempfile.mkdtemp(prefix="wg_calib_")
        with open(os.path.join(d, "CALIB_IN.txt"), "w") as f:
            f.write(f"{y0} {y1} {sc['gamma0']!r} {sc['s0']!r} {sc['s1']!r} {station}\n")
            for i in range(ny):
                f.write(f"{float(base[i])!r} {float(inflow[i])!r} {float(use[i])!r}\n")
        with open(os.path.join(d, "RIVER.DAT"), "w") as f:
            for i in range(ny):
                if y0 + i not in sc["skip_years"]:
                    f.write(f"{y0 + i} {obs_m3s[i]:.2f}\n")
        sbasin.tofile(os.path.join(d, "SBASIN.bin"))
        for i in range(ny):
            pot[i].astype(">f4").tofile(os.path.join(d, f"G_POT_CELL_RUNOFF_{y0 + i}.UNF0"))
        txt = subprocess.run([exe, "calib", d], check=True, capture_output=True, text=True).stdout
        calls = [ln.split()[1:] for ln in txt.splitlines() if ln.startswith("CALIB ")]
        end = [ln.split()[1:] for ln in txt.splitlines() if ln.startswith("CALIB_END ")]
        rd = lambda fn: [ln.rstrip("\n") for ln in open(os.path.join(d, fn)) if not ln.startswith("#")] if os.path.exists(os.path.join(d, fn)) else None
        rec = dict(sc, calls=calls, end=end[0] if end else None, calibration_out=rd("CALIBRATION.OUT"),
                   stat_corr_factor_out=rd("STAT_CORR_FACTOR.OUT"), calibstatus_out=rd("CALIBSTATUS.OUT"))
        g = os.path.join(d, "G_CORR_FACTOR.UNF0")
        rec["corr_factor_grid"] = [float(v) for v in np.fromfile(g, ">f4")] if os.path.exists(g) else None
        out["scenarios"][name] = rec
        print(name, len(calls), "calls", "grid" if rec["corr_factor_grid"] else "", end)
    path = os.path.join(ROOT, "tests", "golden", "ref_calibration.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path) / 1e3, "kB")


if __name__ == "__main__":
    main()
