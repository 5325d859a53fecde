#!/bin/bash
mkdir -p gpurun_out
SHAPES="4k3 4k4 big4" LEGS=sqoa_encode,qoi_encode bash tools/variants.sh 2>&1 | tee gpurun_out/r2n_variants.log
timeout 300 python bench.py --skip-configs --steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for k,v in d['legs'].items(): print(k, round(v['ms'],4),'ms', round(v['frac_of_measured_hbm'],4))
print('value', d['value'], 'e2e', d['e2e']['value'], d['parity_spot_check'])"
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p timeout --timeout=300 2>&1 | tail -2
