#!/bin/bash
# Condense the scratch outputs of `tools/gpu_round.sh <tag>` (gpurun_out/<tag>_*) into the tracked summaries under profiles/.
#   bash tools/make_profiles.sh r1c
TAG=${1:?tag}
G=gpurun_out
C=$(git rev-parse --short HEAD)
for k in k_cells_pre_tpc_m1 k_cells_pre_tpc_m16 k_vertical_tpc_m1 k_vertical_tpc_m16 k_river_level_m1 k_tail_chunk_m1 k_cells_pre_bands_m1 k_days_owner_m1; do
  [ -f $G/${TAG}_$k.ncu-rep ] || continue
  case $k in
    k_days_owner_m1) note="Captured with \`WGK_DAY_SCHEDULE=owner ncu --set full --clock-control none --import-source on -k regex:^k_days_owner\$ -c 1 python tools/profile_run.py --members 1 --days 30 --graph 1\` on a B200 (one launch = 30 simulated days of the whole grid; the opt-in cell-owner schedule); commit $C.";;
    k_cells_pre_bands_m1) note="Captured with \`WGK_VERTICAL_FORM=bands ncu --set full --clock-control none --import-source on -k regex:^k_cells_pre\$ -c 1 python tools/profile_run.py --members 1 --days 1\` on a B200 (band-parallel tile form forced on the full grid); commit $C.";;
    *) kk=${k%_m*}; mm=${k##*_m}; note="Captured with \`ncu --set full --clock-control none --import-source on -k regex:^$kk\$ -c 1 python tools/profile_run.py --members $mm --days 1\` on a B200 (first launch = widest routing level / whole grid); commit $C.";;
  esac
  python tools/ncu_summary.py full $G/${TAG}_$k.ncu-rep profiles/r1_$k.md "$note" | tail -1
done
python tools/ncu_summary.py traffic $G/${TAG}_k_vertical_tpc_m1.ncu-rep profiles/traffic.json 1 | tail -1
python tools/ncu_summary.py traffic $G/${TAG}_k_vertical_tpc_m16.ncu-rep profiles/traffic.json 16 | tail -1
python tools/ncu_summary.py list $G/${TAG}_launches_m1.csv profiles/r1_launches_m1.md "3 simulated days, 1 member, plain launches in wavefront task order (tools/profile_run.py --members 1 --days 3), then one day phase by phase (wgk_profile_day); commit $C."
python tools/ncu_summary.py list $G/${TAG}_launches_m16.csv profiles/r1_launches_m16.md "2 simulated days, 16 members, plain launches in wavefront task order (tools/profile_run.py --members 16 --days 2), then one day phase by phase (wgk_profile_day); commit $C."
python tools/ncu_summary.py list $G/${TAG}_launches_bench.csv profiles/r1_launches_bench.md "first 1500 kernels of \`python bench.py --steps 1 --warmup 1 --no-cpu\` under ncu (set-up, then ~13 simulated days of graph nodes); commit $C."
if [ -f $G/${TAG}_owner_timing.log ]; then
  (echo "# per-warp cycle accounting of the opt-in cell-owner schedule (\`WGK_DAY_SCHEDULE=owner python tools/owner_timing.py --days 365\`, one B200, commit $C)"; echo
   echo 'lane 0 of every warp sums clock64() differences over the four parts of its day; "post" includes the full-mask warp sync in which fast lanes wait for lanes still polling their upstream cells, "vertical+local" includes the CTA barrier at the start of the day.'; echo; echo '```'; cat $G/${TAG}_owner_timing.log; echo '```') > profiles/r1_owner_timing.md
fi
python - "$TAG" "$C" <<'PY'
import json, sys, os
TAG, C = sys.argv[1:3]
G = "gpurun_out/"
def line(fn):
    t = open(G + fn).read().strip().splitlines()
    return t[-1] if t else ""
items = [("python bench.py --steps 30 --warmup 3   (the driver's default line: configs[1], one member)", f"{TAG}_bench_m1.json"),
         ("python bench.py --steps 3 --warmup 3 --members 16 --no-cpu", f"{TAG}_bench_m16.json"),
         ("python bench.py --steps 2 --warmup 3 --members 128 --no-cpu", f"{TAG}_bench_m128.json"),
         ("python bench.py --steps 2 --warmup 3 --members 256 --no-cpu   (configs[3] at its size: 256 members)", f"{TAG}_bench_m256.json"),
         ("python bench.py --steps 1 --warmup 3 --members 1024 --no-cpu   (configs[2] at its size: 1024 parameter sets on ONE GPU)", f"{TAG}_bench_m1024.json"),
         ("python bench.py --steps 2 --warmup 3 --workload 5arcmin --no-cpu   (2 157 440 cells, one member)", f"{TAG}_bench_5arcmin.json"),
         ("WGK_DAY_SCHEDULE=owner python bench.py --steps 5 --warmup 3 --no-cpu   (the opt-in cell-owner schedule: slower, DESIGN.md 4)", f"{TAG}_bench_owner.json"),
         ("python bench.py --impl reference --steps 6 --warmup 3", f"{TAG}_bench_ref.json")]
out, summ = [f"# bench lines of the round-1 box visit (`tools/gpu_round.sh {TAG}`, commit {C}, one B200 unless stated)\n"], []
for cmd, fn in items:
    if not os.path.exists(G + fn) or not line(fn):
        continue
    l = line(fn)
    d = json.loads(l)
    r = d.get("roofline") or {}
    summ.append((cmd.split("   ")[0], d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), r.get("frac"), r.get("step_frac")))
    out.append(f"`{cmd}`\n\n```json\n{l}\n```\n")
out.insert(1, "| command | cell-days/s | ms per step | e2e cell-days/s | roofline.frac (vertical kernel, algorithmic bytes) | step_frac (whole step) |\n|---|---|---|---|---|---|\n"
           + "\n".join(f"| `{c}` | {v:.4g} | {ms:.2f} | {('%.4g' % e) if e else '-'} | {f if f is not None else '-'} | {sf if sf is not None else '-'} |" for c, v, ms, e, f, sf in summ) + "\n")
old = open("profiles/r1_bench.md").read()
i = old.find("`torchrun --nproc-per-node 2 bench.py --gpus 2 --steps 5")
if i >= 0:
    out.append(old[i:])
open("profiles/r1_bench.md", "w").write("\n".join(out))
print("wrote profiles/r1_bench.md")
PY
