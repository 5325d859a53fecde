#!/bin/bash
# tools/gpu_r2_c.sh -- all GPU tests (incl. full-size digests, C tools), encoder timings, bench with the pipelined host path
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -p timeout --timeout=600 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
SHAPES="${SHAPES:-4k3 4k4 big4}" LEGS=sqoa_encode,qoi_encode bash tools/variants.sh 2>&1 | tee gpurun_out/r2c_variants.log
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
for k,v in d['legs'].items(): print(k, round(v['ms'],4),'ms', round(v['mpx_s']),'Mpx/s', round(v['gb_s']),'GB/s', round(v['frac_of_measured_hbm'],4))
print('value', d['value'], 'e2e', d['e2e']['value'], 'parity', d['parity_spot_check'])
PY
SQOA_B200_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_trace.log 2> gpurun_out/bench_trace.err; grep "sqoa_b200\]" gpurun_out/bench_trace.err | tail -12
SQOA_B200_PIPELINE=0 timeout 300 python bench.py --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('no-pipeline e2e', d['e2e']['value'])"
