#!/bin/bash
mkdir -p gpurun_out
VARIANTS="default c76 c92 c124" SHAPES="4k3 4k4 big4" LEGS=sqoa_decode bash tools/variants.sh 2>&1 | tee gpurun_out/r2z_bigchunk.log
