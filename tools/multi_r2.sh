# multi-GPU visit: NCCL moments test (2 GPUs) and the bench line with the sharded legs at N GPUs
#   gpurun --gpus N --timeout 1500 -- 'bash tools/multi_r2.sh N tag'
N=${1:-2}; TAG=${2:-r2m}
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${TAG}_gpus.txt
timeout 600 python -m pytest tests/test_gpu_round2.py tests/test_host_library.py -m gpu -x -q -k "nccl or with_water_use or host_built" > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
tail -c 600 gpurun_out/${TAG}_bench_n$N.json; tail -5 gpurun_out/${TAG}_bench_n$N.err
python - $N $TAG <<'PY'
import json, sys
N, TAG = sys.argv[1:3]
try:
    d = json.loads(open(f"gpurun_out/{TAG}_bench_n{N}.json").read().strip().splitlines()[-1])
    print("N", N, "headline %.4f e9 (%.2f ms/yr)" % (d["value"]/1e9, d["ms_per_step"]))
    for k, v in d.get("sharded", {}).items():
        print("  ", k, "%.4f e9" % (v["value"]/1e9), "ms/step %.2f" % v["ms_per_step"], {x: v.get(x) for x in ("collective_ms_per_step", "gather_ms_per_step", "cells_per_rank", "imbalance_max_over_mean", "beyond_1e-10", "worst_rel", "moments_kernel_equals_host_sum", "error")})
except Exception as e:
    print("no line", e)
PY
