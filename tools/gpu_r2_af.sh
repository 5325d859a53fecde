#!/bin/bash
mkdir -p gpurun_out
for cfg in "1 1 4" "1 1 6" "1 1 8" "1 1 12" "2 2 4" "2 2 6" "2 2 8" "2 2 12" "3 3 6" "3 3 9" "3 3 12" "4 4 8" "4 4 12" "2 2 8" "1 1 8"; do
  set -- $cfg
  SQOA_B200_HOST_CONTEXTS=$2 SQOA_B200_COPY_THREADS=$3 timeout 120 python tools/gpu_e2e_mt.py --threads $1 --images 24 --reps 4 2>&1 | tail -1
done | tee gpurun_out/r2af_e2e_mt.log
