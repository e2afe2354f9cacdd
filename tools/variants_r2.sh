cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; tail -5 gpurun_out/r2e_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2e_smoke.log 2>&1; tail -3 gpurun_out/r2e_smoke.log
# single-member knobs (one simulated year per step)
for K in "WGK_TAIL_THRESHOLD=128" "WGK_TAIL_THRESHOLD=512" "WGK_TAIL_THRESHOLD=1024" "WGK_LEVELS_PER_CHUNK=2" "WGK_LEVELS_PER_CHUNK=3" "X=0"; do
  env $K timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu --legs none > gpurun_out/knob.json 2> gpurun_out/knob.err
  python - "$K" <<'PY'
import json, sys
try:
    d=json.load(open("gpurun_out/knob.json")); g=d["roofline"]["dominant_kernel"]["in_graph"]
    print(sys.argv[1], "%.3f ms/yr, %.4f e9 cd/s, launches/yr %d, V0 %.1f us R0 %.1f us period %.1f us" % (d["ms_per_step"], d["value"]/1e9, d["gpu_launches"]/d["steps"], g["vertical_task_us"], g["river_task_us"], g["day_period_us"]))
except Exception as e: print(sys.argv[1], "failed", e, open("gpurun_out/knob.err").read()[-300:])
PY
done
