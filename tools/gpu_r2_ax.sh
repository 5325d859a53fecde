#!/bin/bash
mkdir -p gpurun_out
for mb in 80 160 400 1200 6000; do
  SQOA_B200_TRANSCODE_GROUP_MB=$mb timeout 300 python bench.py --only cfg5 --steps 5 --warmup 3 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); c=d.get('configs',{}).get('cfg5',d); print('group $mb MB', {k:round(v['ms'],3) for k,v in c['legs'].items()}, round(c['value']), c.get('parity'))
"
done 2>&1 | tee gpurun_out/r2ax_transcode_groups.log
