#!/usr/bin/env python3
"""Condense the bench lines of the final round-2 box visits (gpurun_out/r2f_bench.json, r2f_bench_ref.json, r2g2_bench_n2.json,
r2g8_bench_n8.json) into profiles/r2_bench.md."""
import json
import os
import subprocess

G = "gpurun_out/"
C = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()


def line(fn):
    try:
        return json.loads(open(G + fn).read().strip().splitlines()[-1])
    except Exception:
        return None


runs = [(1, "python bench.py --steps 30 --warmup 3", line("r2f_bench.json")),
        (2, "torchrun --nproc-per-node 2 bench.py --gpus 2 --steps 10 --warmup 3", line("r2g2_bench_n2.json")),
        (8, "torchrun --nproc-per-node 8 bench.py --gpus 8 --steps 10 --warmup 3", line("r2g8_bench_n8.json"))]
out = [f"# bench lines of the final round-2 box visits (commit {C}; B200, 1 965 MHz, no throttle reason)", "",
       "Never taken under a profiler.  `value` = device-timed (CUDA events on the context's stream, max over ranks), `e2e` = through the C ABI "
       "with pinned host buffers.  Every sharded leg carries its parity sample against `oracle/wg_oracle.c` (floors 1e-9 km3 / 1e-6 mm / 1e-6).", "",
       "| N | headline (1 member per GPU) | ms / simulated year | e2e | step_frac (algorithmic bytes / HBM peak) | sweep_1024 | enkf_256 (all-reduce) | basins_5arcmin |",
       "|---|---|---|---|---|---|---|---|"]
for n, cmd, d in runs:
    if not d:
        continue
    s = d.get("sharded", {})
    f = lambda k: f"{s[k]['value'] / 1e9:.2f} x 10^9" if k in s and "value" in s[k] else "-"
    coll = s.get("enkf_256", {}).get("collective_ms_per_step")
    out.append(f"| {n} | {d['value'] / 1e9:.3f} x 10^9 | {d['ms_per_step']:.2f} | {d['e2e']['value'] / 1e9:.3f} x 10^9 | {d['roofline']['step_frac']} | {f('sweep_1024')} | "
               f"{f('enkf_256')} ({coll} ms) | {f('basins_5arcmin')} |")
ref = line("r2f_bench_ref.json")
d1 = runs[0][2]
out += [""]
if ref and d1:
    out += [f"Reference arm on the same box (`python bench.py --impl reference --steps 6 --warmup 3`, the compiled reference, 8 OpenMP threads): "
            f"{ref['value'] / 1e6:.3f} x 10^6 cell-days/s -> e2e ratio {d1['e2e']['value'] / ref['value']:.0f}x, device-timed {d1['value'] / ref['value']:.0f}x.", ""]
if d1:
    r = d1["roofline"]
    ig = r["dominant_kernel"].get("in_graph", {})
    out += [f"Roofline object of the N = 1 line: kernel `{r['dominant_kernel']['name']}` ({r['kernel']}); step {r['step_achieved']} GB/s algorithmic = "
            f"{r['step_frac']} of {r['peak']} GB/s ({r['peak_source']}); traffic of its widest launch {r['traffic']} B (ncu, `profiles/traffic.json`); "
            f"FP64 {r['fp64']['step_tflops']} of {r['fp64']['peak_tflops_measured']} TFLOP/s measured DFMA peak ({r['fp64']['step_frac']}); inside the running graph the level-0 task "
            f"takes {ig.get('vertical_task_us')} us, the day period is {ig.get('day_period_us')} us (median) / {ig.get('day_period_mean_us')} us (mean).", "",
            f"`e2e_classes` (January through the drop-in C++ classes day by day): {d1['e2e_classes']['value'] / 1e6:.2f} x 10^6 cell-days/s; "
            f"`cpu_baseline` sample: {d1['cpu_baseline']['value'] / 1e6:.3f} x 10^6 ({d1['cpu_baseline']['cores']} threads).", ""]
out += ["## raw lines", ""]
for n, cmd, d in runs:
    if d:
        out += [f"`{cmd}`", "", "```json", json.dumps(d), "```", ""]
if ref:
    out += ["`python bench.py --impl reference --steps 6 --warmup 3`", "", "```json", json.dumps(ref), "```", ""]
open("profiles/r2_bench.md", "w").write("\n".join(out))
print("wrote profiles/r2_bench.md")
