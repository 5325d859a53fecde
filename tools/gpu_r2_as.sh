#!/bin/bash
mkdir -p gpurun_out
VARIANTS="default m9" SHAPES="4k3 big3" LEGS=sqoa_decode bash tools/variants.sh 2>&1 | tee gpurun_out/r2as_m9.log
