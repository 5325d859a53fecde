#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
VARIANTS="default r7 r8b" SHAPES="4k3 4k4 big4 1080p4" LEGS=qoi_decode bash tools/variants.sh 2>&1 | tee gpurun_out/r2au_rows8.log
for v in default r7; do
  if [ "$v" = default ]; then L=""; else L=$PWD/gpurun_variants/libsqoa_b200_$v.so; fi
  SQOA_B200_LIB=$L timeout 300 python bench.py --only cfg3 --steps 10 --warmup 3 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); c=d.get('configs',{}).get('cfg3',d); print('$v cfg3', {k:round(v['ms'],3) for k,v in c['legs'].items()}, c.get('parity'))
"
  SQOA_B200_LIB=$L timeout 300 python bench.py --skip-configs --steps 20 --warmup 3 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$v cfg2', round(d['value']), {k:round(v['ms']/16*1000,1) for k,v in d['legs'].items()})
"
done 2>&1 | tee -a gpurun_out/r2au_rows8.log
