#!/bin/bash
# Round-2 GPU-box visits (one section per call; everything lands in gpurun_out/):
#   gpurun --timeout 1500 -- 'bash tools/gpu_r2.sh a r2a'
SEC=${1:-a}
TAG=${2:-r2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/${TAG}_gpu.txt 2>&1
case $SEC in
a)  # first visit: what the tightened floors do to the existing tests, the parity report, a first bench line with the sharded legs
  timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_parity.py > $OUT/${TAG}_pytest_new.log 2>&1
  tail -5 $OUT/${TAG}_pytest_new.log
  timeout 900 python tools/parity_report.py --tag $TAG --write-golden-flips > $OUT/${TAG}_parity.log 2>&1
  tail -12 $OUT/${TAG}_parity.log
  cp tests/golden/gpu_flips_ng1000.json $OUT/${TAG}_gpu_flips_ng1000.json 2>/dev/null
  timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
  tail -c 3000 $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
  timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q > $OUT/${TAG}_pytest_old.log 2>&1
  tail -15 $OUT/${TAG}_pytest_old.log
  ;;
t)  # the whole GPU suite
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
  tail -5 $OUT/${TAG}_pytest.log
  ;;
b)  # bench only
  timeout 900 python bench.py --steps ${STEPS:-30} --warmup 3 ${BENCH_ARGS} > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
  tail -c 3000 $OUT/${TAG}_bench.json; tail -5 $OUT/${TAG}_bench.err
  ;;
p)  # profiles: launch list of the bench command, full captures of the dominant kernels (1 member) and of the vertical kernel in both layouts (64 members)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/${TAG}_launches_bench.csv \
      python bench.py --steps 1 --warmup 1 --no-cpu --legs none > $OUT/${TAG}_bench_under_ncu.log 2>&1
  for K in k_cells_pre_tpc k_river_level k_tail_chunk; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^$K\$" -c 1 -f -o $OUT/${TAG}_${K}_m1 \
        python tools/profile_run.py --members 1 --days 1 > $OUT/${TAG}_ncu_${K}_m1.log 2>&1
  done
  for L in cells members; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^k_vertical_tpc$' -c 1 -f -o $OUT/${TAG}_k_vertical_tpc_m64_$L \
        python tools/profile_run.py --members 64 --days 1 --layout $L > $OUT/${TAG}_ncu_k_vertical_tpc_m64_$L.log 2>&1
    tail -1 $OUT/${TAG}_ncu_k_vertical_tpc_m64_$L.log
  done
  ;;
f)  # final visit: the whole GPU suite, the default bench line, the reference arm, launch list and full capture of the fused task
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
  tail -4 $OUT/${TAG}_pytest.log
  timeout 900 python bench.py --steps 30 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
  tail -c 400 $OUT/${TAG}_bench.err
  timeout 600 python bench.py --impl reference --steps 6 --warmup 3 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
  tail -c 600 $OUT/${TAG}_bench_ref.json
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $OUT/${TAG}_launches_bench.csv \
      python bench.py --steps 1 --warmup 1 --no-cpu --legs none > $OUT/${TAG}_bench_under_ncu.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'^k_level_day$' -c 1 -f -o $OUT/${TAG}_k_level_day_m1 \
      python tools/profile_run.py --members 1 --days 1 > $OUT/${TAG}_ncu_k_level_day_m1.log 2>&1
  tail -2 $OUT/${TAG}_ncu_k_level_day_m1.log
  python smoke_entry.py 2>/dev/null || python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1
  tail -3 $OUT/${TAG}_smoke.log
  ;;
esac
ls -la $OUT | grep ${TAG}_ | tail -20
