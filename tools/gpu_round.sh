#!/bin/bash
# One GPU-box visit: parity tests, the bench lines, the ncu launch list and full captures of the
# dominant kernels.  Everything lands in gpurun_out/ (scratch); tools/ncu_summary.py turns the
# captures into the tracked summaries under profiles/.
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r1b'
TAG=${1:-run}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1

if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
  tail -3 $OUT/${TAG}_pytest.log
fi

# bench lines (never under a profiler)
timeout 600 python bench.py --steps ${STEPS:-10} --warmup 3 > $OUT/${TAG}_bench_m1.json 2> $OUT/${TAG}_bench_m1.err
tail -c 600 $OUT/${TAG}_bench_m1.json
timeout 600 python bench.py --steps 3 --warmup 3 --members 16 --no-cpu > $OUT/${TAG}_bench_m16.json 2> $OUT/${TAG}_bench_m16.err
tail -c 600 $OUT/${TAG}_bench_m16.json
if [ -n "$REF_ARM" ]; then
  timeout 600 python bench.py --impl reference --steps 6 --warmup 3 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
fi

# launch list: every kernel of 3 simulated days, plain launches in wavefront task order
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches_m1.csv \
    python tools/profile_run.py --members 1 --days 3 > $OUT/${TAG}_prof_m1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches_m16.csv \
    python tools/profile_run.py --members 16 --days 2 > $OUT/${TAG}_prof_m16.log 2>&1

# full captures: first launch of each dominant kernel (= the widest level / whole grid)
for K in ${KERNELS:-k_cells_pre k_river_level k_tail_chunk}; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^$K\$" -c 1 -f -o $OUT/${TAG}_${K}_m1 \
      python tools/profile_run.py --members 1 --days 1 > $OUT/${TAG}_ncu_${K}_m1.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^k_cells_pre$' -c 1 -f -o $OUT/${TAG}_k_cells_pre_m16 \
    python tools/profile_run.py --members 16 --days 1 > $OUT/${TAG}_ncu_k_cells_pre_m16.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^k_vertical$' -c 1 -f -o $OUT/${TAG}_k_vertical_m16 \
    python tools/profile_run.py --members 16 --days 1 > $OUT/${TAG}_ncu_k_vertical_m16.log 2>&1
ls -la $OUT | tail -30
