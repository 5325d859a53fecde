#!/bin/bash
mkdir -p gpurun_out
for v in default r9 r10; do
  if [ "$v" = default ]; then L=""; else L=$PWD/gpurun_variants/libsqoa_b200_$v.so; fi
  SQOA_B200_LIB=$L timeout 300 python bench.py --skip-configs --steps 20 --warmup 3 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$v cfg2', round(d['value']), {k:round(v['ms']/16*1000,1) for k,v in d['legs'].items()})
"
  SQOA_B200_LIB=$L timeout 300 python bench.py --only cfg3 --steps 10 --warmup 3 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); c=d.get('configs',{}).get('cfg3',d); print('$v cfg3', {k:round(v['ms'],3) for k,v in c['legs'].items()}, c.get('parity'))
"
done 2>&1 | tee gpurun_out/r2aw_rows910.log
VARIANTS="default r9 r10" SHAPES="big4" LEGS=qoi_decode bash tools/variants.sh 2>&1 | tee -a gpurun_out/r2aw_rows910.log
