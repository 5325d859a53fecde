#!/bin/bash
mkdir -p gpurun_out
VARIANTS="default lw1 lw2 lw8 g1 g2 g4" SHAPES="4k3 4k4" LEGS=sqoa_decode,qoi_decode bash tools/variants.sh 2>&1 | tee gpurun_out/r2x_lanewise.log
