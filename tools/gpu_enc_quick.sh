#!/bin/bash
# tools/gpu_enc_quick.sh -- quick GPU loop for one leg: parity tests, bench, then an ncu capture of the kernel
LEG=${LEG:-sqoa_encode}; KRX=${KRX:-sqoa_encode_block}; TAG=${TAG:-enc}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -p timeout --timeout=150 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
for k,v in d['legs'].items(): print(k, round(v['ms'],4),'ms', round(v['mpx_s']),'Mpx/s', round(v['gb_s']),'GB/s', round(v['frac_of_measured_hbm'],4))
print('e2e', d['e2e']['value'])
PY
python tools/prof_legs.py --legs $LEG --reps 4 > gpurun_out/prof_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$KRX -s 2 -c 2 -o gpurun_out/prof_${TAG} -f python tools/prof_legs.py --legs $LEG --reps 4 > gpurun_out/ncu_${TAG}.log 2>&1; tail -2 gpurun_out/ncu_${TAG}.log
