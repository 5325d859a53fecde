#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "every_decode_attempt or sharded_decode or mixed_qoi" 2>&1 | tail -5
VARIANTS="default c44 c36 c28 c60w896" SHAPES="4k3 big4" LEGS=sqoa_decode bash tools/variants.sh 2>&1 | tee gpurun_out/r2u_chunks.log
