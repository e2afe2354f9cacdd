for v in p0 p1; do STEPS=5 bash tools/variants_bench.sh v9 $v; done
