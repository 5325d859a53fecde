#!/bin/bash
mkdir -p gpurun_out
nproc; lscpu | grep -i "model name\|socket\|numa node(s)" ; cat /sys/kernel/mm/transparent_hugepage/enabled
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p timeout --timeout=300 -k "shard" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for mode in "" "--pinned"; do
  echo "== pipeline on $mode"; SQOA_B200_TRACE=1 timeout 120 python tools/gpu_e2e.py $mode --reps 6 2> gpurun_out/e2e_trace$mode.err; grep "pipeline:" gpurun_out/e2e_trace$mode.err | tail -4
  echo "== pipeline off $mode"; SQOA_B200_PIPELINE=0 timeout 120 python tools/gpu_e2e.py $mode --reps 6
done
echo "== big3 pipeline on"; timeout 120 python tools/gpu_e2e.py --shape big3 --reps 4
echo "== big3 pipeline off"; SQOA_B200_PIPELINE=0 timeout 120 python tools/gpu_e2e.py --shape big3 --reps 4
echo "== copy threads 4"; SQOA_B200_COPY_THREADS=4 timeout 120 python tools/gpu_e2e.py --reps 6
echo "== copy threads 12"; SQOA_B200_COPY_THREADS=12 timeout 120 python tools/gpu_e2e.py --reps 6
