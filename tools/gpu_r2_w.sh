#!/bin/bash
mkdir -p gpurun_out
VARIANTS="e0w4 e0w8 e0w16 e1w8 e1w16" SHAPES="4k3 4k4" LEGS=sqoa_decode,qoi_decode bash tools/variants.sh 2>&1 | tee gpurun_out/r2w_wide.log
