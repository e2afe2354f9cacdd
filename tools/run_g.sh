for v in z0 z1 z0 z1; do STEPS=5 bash tools/variants_bench.sh v21 $v; done
