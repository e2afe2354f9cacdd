cd $GRAFT_REPO_ROOT
for V in libwgk libwgk_pre4; do
 for NM in 32 64; do
  WGK_LIB=$PWD/watergap2_b200/$V.so timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu --legs enkf --enkf-members $NM > gpurun_out/pre_${V}_$NM.json 2> gpurun_out/pre_${V}_$NM.err
  python - <<PY
import json
try:
    e=json.load(open("gpurun_out/pre_${V}_$NM.json"))["sharded"]["enkf_256"]
    print("$V $NM members: %.4f e9 cd/s, %.2f ms/step" % (e["value"]/1e9, e["ms_per_step"]))
except Exception as ex: print("$V $NM failed", ex, open("gpurun_out/pre_${V}_$NM.err").read()[-300:])
PY
 done
done
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; tail -5 gpurun_out/r2d_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2d_smoke.log 2>&1; tail -3 gpurun_out/r2d_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -c 1500 gpurun_out/r2d_bench.json; tail -3 gpurun_out/r2d_bench.err
