#!/bin/bash
# per-kernel timing of the tc2 forward under SCGIB_DBG elimination masks (experiments only)
for d in "$@"; do
  SCGIB_DBG=$d timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['kernels']
print('dbg $d', {n:round(v['ms_per_launch']*1000,1) for n,v in k.items() if 'gin_fwd' in n})
"
done
