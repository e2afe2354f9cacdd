#!/usr/bin/env python3
"""GPU box: parity of the CUDA path against the oracle / the reference golden under the floors of tests/util.py,
reported WITHOUT hiding anything: worst error with and without floor and every value beyond 1e-10 as
(day, field, index, ref, got, rel).  Writes gpurun_out/<tag>_parity.json; `--write-golden-flips` also writes the
committed list tests/golden/gpu_flips_ng1000.json that test_gpu_vs_reference_golden holds the run to."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from tests.util import ParityReport, golden_day  # noqa: E402


def golden_run(form):
    import watergap2_b200 as wg
    os.environ["WGK_VERTICAL_FORM"] = form
    z = np.load(os.path.join(ROOT, "tests", "golden", "ref_ng1000.npz"))
    golden = {k: z[k] for k in z.files}
    ng = int(golden["ng"])
    d0 = golden_day(golden, 0)
    ro = np.zeros(ng, np.int32)
    ro[d0["routing_cell"] - 1] = np.arange(1, ng + 1)
    m = wg.Model(ng)
    m.set_topology(ro, d0["downstream_cell"])
    m.load(d0)
    m.forcing_reserve(31)
    days = [int(d) for d in golden["days"]]
    rep, per_day, curm = ParityReport(), {}, -1
    for sd in range(1, max(days) + 1):
        doy, mon, dom = ((sd - 1) % 365 + 1, 0 if sd <= 31 else 1, sd if sd <= 31 else sd - 31)
        if mon != curm:
            f = {k: golden[f"forcing{mon + 1}/{k}"] for k in ("P", "T", "SW", "LW")}
            m.set_forcing(0, 31, f["P"], f["T"], f["SW"], f["LW"])
            curm = mon
        m.step_days(doy, mon, dom, dom - 1, 1)
        if sd in days:
            r = ParityReport()
            for name, ref in golden_day(golden, sd).items():
                if m.has_field(name) and name != "status_laf_next":
                    r.add(name, ref, m.get(name), tag=sd)
                    rep.add(name, ref, m.get(name), tag=sd)
            per_day[sd] = r.summary()
    m.close()
    return rep, per_day


def free_run(ng, ndays, block):
    from oracle import synth_world as sw, wg_init
    from tests.test_gpu_parity import _run_pair
    w = sw.build_world(ng)
    oracles, m = _run_pair(w, ndays, block=block)
    rep = ParityReport()
    for name in wg_init.STATE_FIELDS + wg_init.FLUX_FIELDS:
        rep.add(name, oracles[0].field(name), m.get(name), tag=ndays)
    m.close()
    return rep


def one_step(ng, ndays, every):
    """both sides restart every day from the oracle's state"""
    from oracle import synth_world as sw, wg_init, wgo
    import watergap2_b200 as wg
    w = sw.build_world(ng)
    ini = wg_init.derive(w)
    topo = ini["_topology"]
    o = wgo.Oracle(w.ng)
    for k, v in ini.items():
        if not k.startswith("_") and o.has(k):
            o.set(k, v)
    m = wg.Model(w.ng)
    m.set_topology(topo["rout_order"], topo["outflow_cell"])
    m.load(ini)
    m.forcing_reserve(31)
    rep = ParityReport()
    for sd in range(1, ndays + 1):
        doy, mon, dom = wgo.calendar(sd)
        if dom == 1:
            f = sw.forcing_month(w, 1901, mon + 1)
            o.set_forcing_month(f)
            m.set_forcing(0, 31, f["P"], f["T"], f["SW"], f["LW"])
        if sd > 1:
            for name in wg_init.STATE_FIELDS + ["storage_transfer"]:
                m.set(name, o.field(name))
        o.step_day(doy, mon, dom)
        m.step_days(doy, mon, dom, dom - 1, 1)
        if sd % every == 0 or sd == ndays:
            for name in wg_init.STATE_FIELDS + wg_init.FLUX_FIELDS:
                rep.add(name, o.field(name), m.get(name), tag=sd)
    m.close()
    return rep


def brief(rep, nflips=40):
    s = rep.summary()
    s["flips"] = [list(f) for f in sorted(rep.flips, key=lambda f: -f[5])[:nflips]]
    by_field = {}
    for f in rep.flips:
        by_field[f[1]] = by_field.get(f[1], 0) + 1
    s["flips_by_field"] = by_field
    return s


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="r2")
    ap.add_argument("--write-golden-flips", action="store_true")
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    out = {}
    for form in ("cells", "bands"):
        rep, per_day = golden_run(form)
        out[f"golden_ng1000_{form}"] = {"total": brief(rep, 200), "per_day": per_day}
        print(f"golden {form}: ", {d: (round(s['worst_rel'], 14), s['beyond_1e-10']) for d, s in per_day.items()}, flush=True)
        if a.write_golden_flips:
            path = os.path.join(ROOT, "tests", "golden", "gpu_flips_ng1000.json")
            prev = json.load(open(path)) if os.path.exists(path) and form != "cells" else {"cells": []}
            known = {(int(f[0]), f[1], int(f[2])) for f in prev.get("entries", [])}
            entries = prev.get("entries", [])
            for f in rep.flips:
                if (int(f[0]), f[1], int(f[2])) not in known:
                    entries.append([int(f[0]), f[1], int(f[2]), f[3], f[4], f[5], form])
            json.dump({"what": "values of the 59-day free run on the 1000-cell golden world that differ from the compiled reference by more than 1e-10 "
                               "(floors of tests/util.py): [day, field, index, reference, gpu, relative error, kernel form]; generated on a B200 by "
                               "tools/parity_report.py --write-golden-flips", "entries": entries}, open(path, "w"), indent=0)
    rep = one_step(3000, 365 if not a.quick else 40, 1)
    out["one_step_3000_365d"] = brief(rep)
    print("one-step 3000 cells:", rep.summary(), flush=True)
    rep = free_run(3000, 30, 7)
    out["free_run_3000_30d"] = brief(rep, 100)
    print("free run 3000 cells 30 d:", rep.summary(), flush=True)
    if not a.quick:
        rep = one_step(67420, 60, 3)
        out["one_step_67420_60d"] = brief(rep)
        print("one-step 67420 cells:", rep.summary(), flush=True)
        rep = free_run(67420, 20, 10)
        out["free_run_67420_20d"] = brief(rep, 100)
        print("free run 67420 cells 20 d:", rep.summary(), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"{a.tag}_parity.json"), "w"), indent=1, default=str)


if __name__ == "__main__":
    main()
