#!/bin/bash
# tools/gpu_e2e_trace.sh -- end-to-end line of bench.py with the host-path phase times, for a few settings
for slots in ${SLOTS:-1 8}; do
echo "== stage slots $slots"
SQOA_B200_STAGE_SLOTS=$slots SQOA_B200_TRACE=1 timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_trace_$slots.log 2> gpurun_out/bench_trace_$slots.err; grep "staged out" gpurun_out/bench_trace_$slots.err | tail -4; python -c "
import json
d=json.loads(open('gpurun_out/bench_trace_$slots.log').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'])"
done
