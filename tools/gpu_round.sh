#!/bin/bash
# One GPU-box visit: parity tests, the bench lines, the ncu launch lists and full captures of the
# dominant kernels.  Everything lands in gpurun_out/ (scratch); tools/ncu_summary.py turns the
# captures into the tracked summaries under profiles/ (see profiles/README.md).
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh r1'
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
timeout 900 python bench.py --steps ${STEPS:-30} --warmup 3 > $OUT/${TAG}_bench_m1.json 2> $OUT/${TAG}_bench_m1.err
tail -c 400 $OUT/${TAG}_bench_m1.json
timeout 600 python bench.py --steps 3 --warmup 3 --members 16 --no-cpu > $OUT/${TAG}_bench_m16.json 2> $OUT/${TAG}_bench_m16.err
timeout 900 python bench.py --steps 2 --warmup 3 --members 128 --no-cpu > $OUT/${TAG}_bench_m128.json 2> $OUT/${TAG}_bench_m128.err
timeout 900 python bench.py --steps 2 --warmup 3 --members 256 --no-cpu > $OUT/${TAG}_bench_m256.json 2> $OUT/${TAG}_bench_m256.err
if [ -n "$BIG" ]; then   # configs[2] at its size: 1024 parameter sets on one GPU (about 2.5 minutes)
  timeout 900 python bench.py --steps 1 --warmup 3 --members 1024 --no-cpu > $OUT/${TAG}_bench_m1024.json 2> $OUT/${TAG}_bench_m1024.err
fi
WGK_DAY_SCHEDULE=owner timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > $OUT/${TAG}_bench_owner.json 2> $OUT/${TAG}_bench_owner.err
timeout 900 python bench.py --steps 2 --warmup 3 --workload 5arcmin --no-cpu > $OUT/${TAG}_bench_5arcmin.json 2> $OUT/${TAG}_bench_5arcmin.err
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -1 $OUT/${TAG}_smoke.log
if [ -n "$REF_ARM" ]; then
  timeout 600 python bench.py --impl reference --steps 6 --warmup 3 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
fi

# launch lists: every kernel of 3 (2) simulated days, plain launches in wavefront task order, then one day
# phase by phase (wgk_profile_day)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches_m1.csv \
    python tools/profile_run.py --members 1 --days 3 > $OUT/${TAG}_prof_m1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches_m16.csv \
    python tools/profile_run.py --members 16 --days 2 > $OUT/${TAG}_prof_m16.log 2>&1

# the same pass over the bench command itself (first 1500 kernels: set-up, then ~13 simulated days of graph nodes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/${TAG}_launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu > $OUT/${TAG}_bench_under_ncu.log 2>&1

# full captures: first launch of each dominant kernel (= the widest level / whole grid)
for K in k_cells_pre_tpc k_river_level k_tail_chunk k_vertical_tpc; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^$K\$" -c 1 -f -o $OUT/${TAG}_${K}_m1 \
      python tools/profile_run.py --members 1 --days 1 > $OUT/${TAG}_ncu_${K}_m1.log 2>&1
done
for K in k_cells_pre_tpc k_vertical_tpc; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^$K\$" -c 1 -f -o $OUT/${TAG}_${K}_m16 \
      python tools/profile_run.py --members 16 --days 1 > $OUT/${TAG}_ncu_${K}_m16.log 2>&1
done
# the band-parallel tile form (small problems), forced on the full grid for comparison
WGK_VERTICAL_FORM=bands timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^k_cells_pre$' -c 1 -f \
    -o $OUT/${TAG}_k_cells_pre_bands_m1 python tools/profile_run.py --members 1 --days 1 > $OUT/${TAG}_ncu_bands_m1.log 2>&1
# the opt-in cell-owner schedule: one launch of 30 simulated days
WGK_DAY_SCHEDULE=owner timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^k_days_owner$' -c 1 -f \
    -o $OUT/${TAG}_k_days_owner_m1 python tools/profile_run.py --members 1 --days 30 --graph 1 > $OUT/${TAG}_ncu_owner_m1.log 2>&1
WGK_DAY_SCHEDULE=owner timeout 300 python tools/owner_timing.py --days 365 > $OUT/${TAG}_owner_timing.log 2>&1
ls -la $OUT | grep ${TAG}_ | tail -40
