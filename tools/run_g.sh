for cfg in "w1 WGK_CLASS_ORDER=asc" "w1 WGK_CLASS_ORDER=cost" "w2 WGK_CLASS_ORDER=asc" "w2 WGK_CLASS_ORDER=cost"; do
  set -- $cfg; v=$1; shift
  echo -n "$* : "; VENV="$*" STEPS=5 bash tools/variants_bench.sh v18 $v
done
