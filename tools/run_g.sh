for M in 2 4 8; do for mode in fused split; do
WGK_LEVEL_TASKS=$mode timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu --legs none --members $M > gpurun_out/r2u_${mode}_$M.json 2>/dev/null
python - $mode $M gpurun_out/r2u_${mode}_$M.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[3]).read().strip().splitlines()[-1])
print(sys.argv[1], 'members', sys.argv[2], '%.3e' % d['value'], round(d['ms_per_step'],2), 'ms/yr')
PY
done; done
