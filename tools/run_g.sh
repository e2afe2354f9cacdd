for v in k0 kf1 kf2 kf1n3 k0n3; do
cp variants/libwgk_$v.so watergap2_b200/libwgk.so
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu --legs none --members 128 > gpurun_out/r2v_$v.json 2>/dev/null
python - $v gpurun_out/r2v_$v.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
print(sys.argv[1], 'members 128', '%.3e' % d['value'], round(d['ms_per_step'],2), 'ms/yr', d['roofline']['kernel'][:100])
PY
done
