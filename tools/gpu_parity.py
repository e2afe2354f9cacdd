#!/usr/bin/env python3
"""GPU-vs-oracle parity run on a synthetic world (development tool; the same comparison
lives in tests/test_gpu_parity.py).  Prints, per compared field, the number of cells whose
relative difference exceeds the tolerance and the worst cell."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import synth_world as sw, wg_init, wgo  # noqa: E402
import watergap2_b200 as wg  # noqa: E402

COMPARE = wg_init.STATE_FIELDS + wg_init.FLUX_FIELDS


def rel_diff(a, b, floor):
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    return np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ng", type=int, default=3000)
    ap.add_argument("--days", type=int, default=59)
    ap.add_argument("--tol", type=float, default=1e-10)
    ap.add_argument("--check-every", type=int, default=1)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--stop-first", action="store_true", help="stop at the first day out of tolerance and dump the worst cell")
    ap.add_argument("--tail-threshold", type=int, default=0)
    a = ap.parse_args()
    w = sw.build_world(a.ng)
    init = wg_init.derive(w)
    topo = init["_topology"]
    o = wgo.Oracle(a.ng)
    for k, v in init.items():
        if not k.startswith("_") and o.has(k):
            o.set(k, v)
    m = wg.Model(a.ng, use_graph=0 if a.no_graph else 1, tail_threshold=a.tail_threshold)
    m.set_topology(topo["rout_order"], topo["outflow_cell"])
    print("levels", m.nlevels, "fields loaded", m.load(init))
    m.forcing_reserve(31)
    worst = {}
    curm = -1
    t_gpu = 0.0
    for sd in range(1, a.days + 1):
        doy, mon, dom = wgo.calendar(sd)
        if mon != curm:
            f = sw.forcing_month(w, 1901, mon + 1)
            o.set_forcing_month(f)
            m.set_forcing(0, 31, f["P"], f["T"], f["SW"], f["LW"])
            curm = mon
        o.step_day(doy, mon, dom)
        t0 = time.time()
        m.step_days(doy, mon, dom, dom - 1, 1)
        m.synchronize()
        t_gpu += time.time() - t0
        if sd % a.check_every and sd != a.days:
            continue
        nbad = 0
        for name in COMPARE:
            x = o.field(name)
            y = m.get(name)
            if x.dtype.kind != "f":
                nb = int((x != y).sum())
                if nb:
                    print(f"  day {sd} {name}: {nb} integer mismatches")
                    nbad += nb
                continue
            from tests.util import floor_of
            d = rel_diff(x, y, floor_of(name))
            nb = int((d > a.tol).sum())
            k = int(np.argmax(d))
            if d[k] > worst.get(name, (0,))[0]:
                worst[name] = (float(d[k]), sd, k, float(x[k]), float(y[k]))
            nbad += nb
            if nb:
                print(f"  day {sd} {name}: {nb} cells > {a.tol:g}; worst cell {k}: oracle {x[k]!r} gpu {y[k]!r}")
        if nbad and a.stop_first:
            first = None
            for name in COMPARE:
                x = o.field(name); y = m.get(name)
                if x.dtype.kind == "f":
                    from tests.util import floor_of
                    d = rel_diff(x, y, floor_of(name))
                    if (d > a.tol).any():
                        first = (name, int(np.argmax(d)))
                        break
            name, k = first
            cell = k // 101 if name == "snow_bands" else k
            print(f"FIRST OFFENDER day {sd}: field {name} index {k} -> cell {cell}; level {m.levels()[cell]}")
            for nm in COMPARE + ["t_inflow_local", "t_runoff_to_river", "t_gw_to_river"]:
                if nm == "snow_bands":
                    continue
                y = m.get(nm)
                x = o.field(nm) if o.has(nm) else None
                print(f"   {nm:24s} oracle {x[cell] if x is not None else None!r:>26} gpu {y[cell]!r}")
            for nm in ["arid", "ldd", "landcover", "loc_lake", "loc_wetland", "glo_wetland", "lake_area", "reservoir_area", "contfreq", "smax", "downstream_cell"]:
                print(f"   static {nm:18s} {np.asarray(init[nm]).ravel()[cell]!r}")
            up = np.nonzero(np.asarray(init["downstream_cell"]) == cell + 1)[0]
            print("   upstream cells", up.tolist(), "discharge oracle", o.field("discharge")[up].tolist(), "gpu", m.get("discharge")[up].tolist())
            break
        ts_o, ts_g = o.total_storage_km3(), m.total_storage_km3()
        print(f"day {sd}: cells out of tolerance {nbad}; total storage oracle {ts_o:.9e} gpu {ts_g:.9e} rel {abs(ts_o-ts_g)/abs(ts_o):.2e}")
    print("worst relative differences:")
    for name, v in sorted(worst.items(), key=lambda kv: -kv[1][0])[:12]:
        print(f"  {name:24s} {v[0]:.3e} (day {v[1]}, cell {v[2]}, oracle {v[3]!r}, gpu {v[4]!r})")
    print(f"gpu wall {t_gpu:.3f}s for {a.days} days ({a.ng*a.days/t_gpu:.3e} cell-days/s incl. launch+sync), launches {m.kernel_launches}")


if __name__ == "__main__":
    main()
