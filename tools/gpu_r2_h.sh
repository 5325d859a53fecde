#!/bin/bash
mkdir -p gpurun_out
make -s -C tools
echo "== sqoabench no reference"; timeout 120 tools/bin/sqoabench_b200 2 --synth cfg1 2>&1 | tail -25; echo "rc=$?"
echo "== sqoabench with reference"; timeout 120 tools/bin/sqoabench_b200 2 --synth cfg1 --reference oracle/_ref/libsqoa_ref.so 2>&1 | tail -25; echo "rc=$?"
timeout 900 python -m pytest tests/test_tools.py tests/test_gpu_full_size.py -q -m gpu -p timeout --timeout=600 --timeout-method=thread 2>&1 | tail -8
SHAPES="${SHAPES:-4k3 4k4 big4}" LEGS=sqoa_encode,qoi_encode bash tools/variants.sh 2>&1 | tee gpurun_out/r2h_variants.log
timeout 600 python bench.py --only cfg3 --steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for k,v in d['legs'].items(): print('cfg3',k, round(v['ms'],3),'ms', round(v['frac_of_measured_hbm'],4))
print('parity', d['parity'])"
timeout 600 python bench.py --only cfg5 --steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for k,v in d['legs'].items(): print('cfg5',k, round(v['ms'],3),'ms')
print('parity', d['parity'], 'value', d['value'])"
