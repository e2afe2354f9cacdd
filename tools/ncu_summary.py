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
    # where the warps spend their time: warp-state samples and executed instructions per device function and per
    # source line (needs -lineinfo and --import-source on); the CUDA+SASS source page carries the metrics
    try:
        import os
        import re
        src = ncu_csv(rep, "source", ["--print-source", "cuda,sass"])
        hdr_i = next(i for i, r in enumerate(src) if r and r[0] == "Line No")
        h = src[hdr_i]
        fpath = next((r[1] for r in src[:hdr_i] if r and r[0] == "File Path"), None)
        stall = [i for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
        num = lambda x: float(x) if x not in ("", "-") else 0.0
        agg = defaultdict(lambda: [0.0, 0.0, "", defaultdict(float)])
        for r in src[hdr_i + 1:]:
            if r and r[0] != "":
                try:
                    ln = int(r[0])
                except ValueError:
                    continue
                agg[ln][0] += num(r[6])
                agg[ln][1] += num(r[7])
                agg[ln][2] = r[1]
                for i in stall:
                    agg[ln][3][h[i][6:]] += num(r[i])
        tot_s = sum(v[0] for v in agg.values()) or 1.0
        tot_i = sum(v[1] for v in agg.values()) or 1.0
        starts = []
        if fpath and os.path.exists(fpath):
            for n, l in enumerate(open(fpath).read().split("\n"), 1):
                mm = re.match(r"^__device__ .*? (\w+)\(", l) or re.match(r"^__global__ void (?:__launch_bounds__\([^)]*\) )?(\w+)\(", l)
                if mm:
                    starts.append((n, mm.group(1)))
        def region(ln):
            name = "(kernel body / inlined libm)"
            for n, f in starts:
                if n <= ln:
                    name = f
                else:
                    break
            return name
        reg = defaultdict(lambda: [0.0, 0.0, defaultdict(float)])
        for ln, v in agg.items():
            r_ = reg[region(ln)]
            r_[0] += v[0]
            r_[1] += v[1]
            for k, x in v[3].items():
                r_[2][k] += x
        lines += ["## warp-state samples and executed warp instructions per device function", "",
                  "(function = the last `__device__`/`__global__` definition above the source line in the current tree)", "",
                  "| function | % samples | % instructions | dominant stall reasons |", "|---|---|---|---|"]
        for k, v in sorted(reg.items(), key=lambda kv: -kv[1][0])[:14]:
            top = sorted(v[2].items(), key=lambda kv: -kv[1])[:3]
            lines.append(f"| {k} | {100*v[0]/tot_s:.1f} | {100*v[1]/tot_i:.1f} | " + ", ".join(f"{a_} {100*b_/max(v[0],1):.0f}%" for a_, b_ in top) + " |")
        lines += ["", "## hottest source lines", "", "| % samples | % instr | line | source | stalls |", "|---|---|---|---|---|"]
        for ln, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:20]:
            top = sorted(v[3].items(), key=lambda kv: -kv[1])[:2]
            lines.append(f"| {100*v[0]/tot_s:.1f} | {100*v[1]/tot_i:.1f} | {ln} | `{v[2][:90].replace('|', '/')}` | " + ", ".join(f"{a_} {100*b_/max(v[0],1):.0f}%" for a_, b_ in top) + " |")
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


def traffic(rep, dst, members):
    """add the DRAM bytes per launch of the kernels in `rep` to the json `dst` ({kernel: {members: bytes}})"""
    import json
    import os
    rows = ncu_csv(rep, "raw")
    hdr, units = rows[0], rows[1]
    out = json.load(open(dst)) if os.path.exists(dst) else {}
    scale = {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "tbyte": 1e12}
    for r in rows[2:]:
        d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
        b = sum(float(d[k]) * scale[u[k].lower()] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        name = d["Kernel Name"].split("(")[0].split("<")[0]
        out.setdefault(name, {})[str(members)] = int(b)
    json.dump(out, open(dst, "w"), indent=1, sort_keys=True)
    print("wrote", dst, out)


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    note = " ".join(sys.argv[4:])
    if mode == "traffic":
        traffic(src, dst, int(sys.argv[4]))
    else:
        (full if mode == "full" else launch_list)(src, dst, note)
