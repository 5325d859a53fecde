#!/bin/bash
mkdir -p gpurun_out
nproc
for cfg in "1 1 16" "1 1 8" "2 2 16" "2 2 32" "3 3 24" "4 4 32" "4 2 16" "2 2 8" "16 2 16" "16 4 32"; do
  set -- $cfg
  SQOA_B200_HOST_CONTEXTS=$2 SQOA_B200_COPY_THREADS=$3 timeout 120 python tools/gpu_e2e_mt.py --threads $1 --images 16 2>&1 | tail -1
done | tee gpurun_out/r2ae_e2e_mt.log
