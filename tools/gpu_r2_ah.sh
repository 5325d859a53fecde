#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -4
for l in 1 0; do
echo "== SQOA_B200_QOI_LANES=$l"
SQOA_B200_QOI_LANES=$l VARIANTS="default" SHAPES="4k3 big3" LEGS=qoi_decode bash tools/variants.sh 2>&1
done | tee gpurun_out/r2ah_lanes.log
