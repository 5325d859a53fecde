#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02_reference_n1.json 2> gpurun_out/r02_reference_n1.err
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
echo "rc=$?"; tail -2 gpurun_out/r02_bench_n1.err
make -C tools > /dev/null 2>&1; timeout 300 tools/bin/sqoabench_b200 5 --synth cfg1 --synth cfg2 --reference oracle/_ref/libsqoa_ref.so > gpurun_out/r02_sqoabench.txt 2>&1; tail -12 gpurun_out/r02_sqoabench.txt
