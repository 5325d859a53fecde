#!/bin/bash
mkdir -p gpurun_out
timeout 40 python bench.py --skip-configs --steps 20 --warmup 3 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('final default', round(d['value']), {k:round(v['ms']/16*1000,1) for k,v in d['legs'].items()}, d.get('parity_spot_check'), round(d['e2e']['value']))
" | tee gpurun_out/r2bb_final.log
