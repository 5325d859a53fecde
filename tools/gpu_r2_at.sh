#!/bin/bash
mkdir -p gpurun_out
for v in default m9; do
  if [ "$v" = default ]; then L=""; else L=$PWD/gpurun_variants/libsqoa_b200_$v.so; fi
  SQOA_B200_LIB=$L timeout 300 python bench.py --skip-configs --steps 20 --warmup 3 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$v', round(d['value']), {k:round(v['ms']/16*1000,1) for k,v in d['legs'].items()})
"
done | tee gpurun_out/r2at_m9_bench.log
