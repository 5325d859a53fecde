#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
VARIANTS="default" SHAPES="4k3 4k4 big4 1080p4" LEGS=sqoa_decode bash tools/variants.sh 2>&1 | tee gpurun_out/r2al_valprefetch.log
