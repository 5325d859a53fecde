#!/bin/bash
# experiment: how much of the QOI encoder's time is the match instruction of the index rows?
mkdir -p gpurun_out
VARIANTS="default nomatch" SHAPES="4k3 big4" LEGS=qoi_encode bash tools/variants.sh 2>&1 | tee gpurun_out/r2q_nomatch.log
