#!/bin/bash
mkdir -p gpurun_out
make -s -C tools
echo "== sqoabench"; timeout 120 tools/bin/sqoabench_b200 2 --synth cfg1 > gpurun_out/sqoabench.out 2> gpurun_out/sqoabench.err; echo "rc=$?"; tail -12 gpurun_out/sqoabench.out; tail -30 gpurun_out/sqoabench.err
echo "== decoder occupancy variants"
VARIANTS="default d6 d8" SHAPES="4k3 big4" LEGS=sqoa_decode,qoi_decode bash tools/variants.sh 2>&1 | tee gpurun_out/r2i_variants.log
