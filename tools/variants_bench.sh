#!/bin/bash
# development tool: headline bench (one member, 365-day graph) for kernel variants built by tools/mkvariant.sh
#   gpurun -- 'bash tools/variants_bench.sh tag name1 name2 ...'   (extra environment: VENV="WGK_X=1 ...")
tag=$1; shift
mkdir -p gpurun_out
[ -n "$GRAFT_REPO_ROOT" ] || { echo "run on the GPU box (the variant replaces watergap2_b200/libwgk.so of the scratch copy)"; exit 1; }
for v in "$@"; do
  # libwghost.so links watergap2_b200/libwgk.so: the variant must BE that file, or two copies of the library get mixed
  cp variants/libwgk_$v.so watergap2_b200/libwgk.so
  env $VENV timeout 300 python bench.py --steps ${STEPS:-6} --warmup 3 --no-cpu --legs none > gpurun_out/${tag}_$v.json 2> gpurun_out/${tag}_$v.err
  python - "$v" gpurun_out/${tag}_$v.json <<'PY'
import sys, json
v, f = sys.argv[1:3]
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    g = d["roofline"]["dominant_kernel"].get("in_graph", {})
    print(f"{v:16s} {d['ms_per_step']:7.2f} ms/yr  e2e {d['e2e']['value']:.3e}  V0 {g.get('vertical_task_us')} R0 {g.get('river_task_us')} period {g.get('day_period_us')} us", flush=True)
except Exception as e:
    print(v, "FAILED", e)
PY
done
