#!/bin/bash
mkdir -p gpurun_out
for b in 3 2 1; do
  echo "== blocks per SM $b"; SQOA_B200_ENC_BLOCKS_PER_SM=$b SHAPES="4k3 big4" LEGS=sqoa_encode,qoi_encode bash tools/variants.sh 2>&1
done | tee gpurun_out/r2o_grid.log
