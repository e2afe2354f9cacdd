#!/usr/bin/env python3
"""Development tool: a simulated year of the 0.5 degree world with M members under each schedule of a multi-day call
(WGK_DAY_SCHEDULE = wavefront | wholeday) and as 365 three-call days (wgk_vertical_day + wgk_routing_day, plain launches),
plus the per-month time of the default schedule (seasonality of the band-loop skip)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--members", type=int, default=64)
    a = ap.parse_args()
    w, ini = bench.build_inputs()
    forcing = bench.year_forcing(w)

    def year(m, call):
        for _ in range(2):
            call(m)
        m.synchronize()
        t0 = time.perf_counter()
        call(m)
        m.synchronize()
        return (time.perf_counter() - t0) * 1e3

    for sched in ("wavefront", "wholeday"):
        os.environ["WGK_DAY_SCHEDULE"] = sched
        m = bench.make_model(w, ini, a.members, 0)
        bench.upload_year(m, forcing)
        ms = year(m, lambda m: m.step_days(1, 0, 1, 0, 365))
        print(f"members {a.members} {sched}: {ms:.1f} ms/yr = {ms / 365 / a.members * 1e3:.1f} us per member-day")
        if sched == "wholeday":
            def months(m):
                doy = 1
                for mon, nd in enumerate(bench.NDAYS):
                    m.synchronize()
                    t0 = time.perf_counter()
                    m.step_days(doy, mon, 1, doy - 1, nd)
                    m.synchronize()
                    per.append((time.perf_counter() - t0) * 1e3 / nd)
                    doy += nd
            per = []
            months(m)
            per = []
            months(m)
            print("  ms per day by month:", " ".join(f"{x:.2f}" for x in per))

            def shim(m):
                doy = 1
                for mon, nd in enumerate(bench.NDAYS):
                    for d in range(nd):
                        m.vertical_day(doy, mon, d + 1, doy - 1)
                        m.routing_day(doy, mon, d + 1)
                        doy += 1
            ms = year(m, shim)
            print(f"members {a.members} three-call days (plain launches): {ms:.1f} ms/yr = {ms / 365 / a.members * 1e3:.1f} us per member-day")
        m.close()


if __name__ == "__main__":
    main()
