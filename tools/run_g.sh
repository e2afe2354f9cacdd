cp variants/libwgk_x1.so watergap2_b200/libwgk.so
for o in asc cost; do
WGK_CLASS_ORDER=$o timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu --legs sweep,enkf > gpurun_out/r2z_$o.json 2>/dev/null
python - $o gpurun_out/r2z_$o.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
print(sys.argv[1], {k: '%.4e' % v['value'] for k,v in d['sharded'].items()})
PY
done
