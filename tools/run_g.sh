for o in asc desc cost; do echo -n "$o: "; VENV="WGK_CLASS_ORDER=$o" STEPS=4 bash tools/variants_bench.sh v7_$o j128; done
