for cfg in "s4 WGK_CLASS_ORDER=cost" "s3 WGK_CLASS_ORDER=cost" "r128 WGK_CLASS_ORDER=cost" "s4 WGK_CLASS_ORDER=asc"; do
  set -- $cfg; v=$1; shift
  echo -n "$* : "; VENV="$*" STEPS=5 bash tools/variants_bench.sh v14 $v
done
