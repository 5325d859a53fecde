#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p timeout --timeout=300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for mode in "" "--pinned"; do
  echo "== pipeline on $mode"; SQOA_B200_TRACE=1 timeout 120 python tools/gpu_e2e.py $mode --reps 8 2> gpurun_out/e2e_trace$mode.err; grep "pipeline:" gpurun_out/e2e_trace$mode.err | tail -4
done
echo "== pipeline off"; SQOA_B200_PIPELINE=0 timeout 120 python tools/gpu_e2e.py --reps 8
echo "== big3 pipeline on"; timeout 120 python tools/gpu_e2e.py --shape big3 --reps 4
echo "== 1080p4 on"; timeout 120 python tools/gpu_e2e.py --shape 1080p4 --reps 8
echo "== 1080p4 off"; SQOA_B200_PIPELINE=0 timeout 120 python tools/gpu_e2e.py --shape 1080p4 --reps 8
