#!/bin/bash
mkdir -p gpurun_out
VARIANTS="default rows5p rows6 rows6w" SHAPES="4k3 4k4 big4" LEGS=qoi_decode bash tools/variants.sh 2>&1 | tee gpurun_out/r2ap_rows_occ.log
