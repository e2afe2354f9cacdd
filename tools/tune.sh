#!/bin/bash
# sensitivity of the single-member run to the task granularity knobs (development tool)
for cfg in "256 1 0" "256 1 1" "1024 1 1" "64 1 1" "4096 1 1"; do
  set -- $cfg
  WGK_TAIL_THRESHOLD=$1 WGK_LEVELS_PER_CHUNK=$2 WGK_FUSE_NARROW=$3 python bench.py --steps 3 --warmup 2 --no-cpu --members ${MEMBERS:-1} 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tail_threshold $1 levels_per_chunk $2 fuse_narrow $3:', round(d['ms_per_step'],2), 'ms/yr', d['gpu_launches'], 'launches', '%.3e' % d['value'], 'e2e %.3e' % d['e2e']['value'])"
done
