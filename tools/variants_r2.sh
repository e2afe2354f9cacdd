cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -s -k "water_use" > gpurun_out/r2f_wu.log 2>&1; tail -25 gpurun_out/r2f_wu.log | cut -c1-1200
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; tail -4 gpurun_out/r2f_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --legs none > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2f_bench.json")); print("headline %.3f ms/yr %.4f e9" % (d["ms_per_step"], d["value"]/1e9))
PY
