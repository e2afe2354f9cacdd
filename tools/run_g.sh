for v in u0 w1; do STEPS=5 bash tools/variants_bench.sh v17 $v; done
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -x -q -m gpu 2>&1 | tail -3
