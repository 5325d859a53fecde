#!/bin/bash
mkdir -p gpurun_out
VARIANTS="default eager" SHAPES="4k3 big4" LEGS=sqoa_decode,qoi_decode bash tools/variants.sh 2>&1 | tee gpurun_out/r2l_variants.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p timeout --timeout=300 2>&1 | tail -3
