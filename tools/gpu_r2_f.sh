#!/bin/bash
mkdir -p gpurun_out
echo "== pipeline on"; SQOA_B200_TRACE=1 timeout 120 python tools/gpu_e2e.py --reps 6 2> gpurun_out/e2e_trace.err; grep "pipeline:\|progress of" gpurun_out/e2e_trace.err | tail -8
