for cfg in "t0 WGK_LEVEL_TASKS=fusedfree" "t0 WGK_LEVEL_TASKS=fused"; do
  set -- $cfg; v=$1; shift
  echo -n "$* : "; VENV="$*" STEPS=5 bash tools/variants_bench.sh v15 $v
done
cp variants/libwgk_t0.so watergap2_b200/libwgk.so
WGK_LEVEL_TASKS=fusedfree LEVELS=0,1,3,8,20,40,56 python tools/level_timeline.py 2>&1 | tail -8
