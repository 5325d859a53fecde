#!/bin/bash
mkdir -p gpurun_out
VARIANTS="default norows norows_notab" SHAPES="4k3 big4" LEGS=qoi_encode bash tools/variants.sh 2>&1 | tee gpurun_out/r2s_exp.log
