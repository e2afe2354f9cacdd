#!/usr/bin/env python3
"""Turn ncu outputs (gpurun_out/, scratch) into the tracked summaries under profiles/.

  python tools/ncu_summary.py full   gpurun_out/x.ncu-rep  profiles/r1_x.md      # one --set full capture
  python tools/ncu_summary.py list   gpurun_out/launches.csv profiles/r1_launches.md [--days N]

`full` keeps the counters that decide the roofline discussion (time, DRAM bytes, achieved
bandwidth, FP64 pipe, issue slots, occupancy and its limiter, stall reasons, L1/L2 hit rates) and the
hottest source lines.  `list` aggregates a `--metrics gpu__time_duration.sum` launch list per kernel.
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict, defaultdict

KEEP = [
    ("gpu__time_duration.sum", "kernel time"),
    ("launch__grid_size", "grid size (CTAs)"),
    ("launch__block_size", "block size"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_static", "static smem / CTA"),
    ("launch__occupancy_limit_registers", "occupancy limit: registers (CTAs/SM)"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit: smem (CTAs/SM)"),
    ("launch__occupancy_limit_warps", "occupancy limit: warps (CTAs/SM)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__bytes.sum.per_second", "DRAM bandwidth"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (% of ncu peak)"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe instructions (% of peak)"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp instruction"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (global/local memory)"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait (fixed latency)"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard (smem/MUFU)"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall: branch resolving"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall: LG throttle"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall: no instruction"),
]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def full(rep, dst, note=""):
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full summary: `{rep.split('/')[-1]}`", ""]
    if note:
        lines += [note, ""]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        lines += [f"## {d.get('Kernel Name', '?')}  grid {d.get('Grid Size', '?')} block {d.get('Block Size', '?')}", "",
                  "| counter | value | unit |", "|---|---|---|"]
        for k, label in KEEP:
            if k in d and d[k] != "":
                lines.append(f"| {label} (`{k}`) | {d[k]} | {u[k]} |")
        try:
            def gb(key):
                v, un = float(d[key]), u[key].lower()
                return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "tbyte": 1e12}[un]
            tr = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
            t = float(d["gpu__time_duration.sum"]) * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}[u["gpu__time_duration.sum"]]
            lines += ["", f"traffic = DRAM read + write = {tr/1e6:.2f} MB per launch; {tr/t/1e9:.0f} GB/s under ncu (cold cache, serialised)."]
        except Exception as e:  # noqa: BLE001
            lines += ["", f"(traffic not derived: {e})"]
        lines.append("")
    # hottest source lines (needs -lineinfo and --import-source on)
    try:
        src = ncu_csv(rep, "source", ["--print-source", "cuda"])
        if src and len(src) > 2:
            h = src[0]
            def col(name):
                for i, x in enumerate(h):
                    if x.strip() == name:
                        return i
                return None
            ci, cs, cl = col("# Samples") or col("Sampling Data (All)"), col("Source"), col("#")
            if ci is None:
                for i, x in enumerate(h):
                    if "Samples" in x or "Sampling" in x:
                        ci = i
                        break
            if ci is not None and cs is not None:
                scored = []
                for r in src[1:]:
                    try:
                        scored.append((float(r[ci]), r[cl] if cl is not None else "", r[cs].strip()))
                    except Exception:  # noqa: BLE001
                        pass
                tot = sum(s for s, _, _ in scored) or 1.0
                scored.sort(reverse=True)
                lines += ["## hottest source lines (warp-state samples)", "", "| % samples | line | source |", "|---|---|---|"]
                for s, ln, text in scored[:25]:
                    lines.append(f"| {100*s/tot:.1f} | {ln} | `{text[:110].replace('|', '/')}` |")
                lines.append("")
    except Exception as e:  # noqa: BLE001
        lines += [f"(source page not available: {e})", ""]
    open(dst, "w").write("\n".join(lines))
    print("wrote", dst)


def launch_list(path, dst, note=""):
    rows = []
    with open(path, newline="") as fh:
        text = fh.read()
    start = text.find('"ID"')
    rd = csv.DictReader(io.StringIO(text[start:]))
    per = OrderedDict()
    tot = 0.0
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        un = r["Metric Unit"]
        ns = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6, "second": 1e9}[un]
        name = r["Kernel Name"].split("(")[0]
        k = per.setdefault(name, [0, 0.0, 0.0])
        k[0] += 1
        k[1] += ns
        k[2] = max(k[2], ns)
        tot += ns
        rows.append((name, ns))
    lines = [f"# ncu launch list: `{path.split('/')[-1]}`", "", note, "",
             "per-launch times under ncu are cold-cache and serialised: the SHARE of the step is what carries over.", "",
             "| kernel | launches | total µs | share | mean µs | max µs |", "|---|---|---|---|---|---|"]
    for name, (n, ns, mx) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| {name} | {n} | {ns/1e3:.1f} | {100*ns/tot:.1f} % | {ns/n/1e3:.2f} | {mx/1e3:.1f} |")
    lines += ["", f"total {tot/1e3:.1f} µs over {len(rows)} launches", ""]
    open(dst, "w").write("\n".join(lines))
    print("wrote", dst)


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    note = " ".join(sys.argv[4:])
    (full if mode == "full" else launch_list)(src, dst, note)
