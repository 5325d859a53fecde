#!/bin/bash
mkdir -p gpurun_out
VARIANTS="r6w r6a r6b r6c" SHAPES="4k3 4k4 big4 1080p4" LEGS=qoi_decode bash tools/variants.sh 2>&1 | tee gpurun_out/r2aq_rows_occ.log
for v in r6w r6a; do echo "== cfg3/cfg5 with $v"; SQOA_B200_LIB=$PWD/gpurun_variants/libsqoa_b200_$v.so timeout 300 python bench.py --only cfg3 --steps 10 --warmup 3 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); c=d.get('configs',{}).get('cfg3',d); print({k:round(v['ms'],3) for k,v in c['legs'].items()}, c.get('parity'))
"; done 2>&1 | tee -a gpurun_out/r2aq_rows_occ.log
