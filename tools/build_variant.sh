#!/bin/bash
# tools/build_variant.sh <name> <extra nvcc -D flags...> -- builds gpurun_out/variants/libsqoa_b200_<name>.so for tuning runs
set -e
name=$1; shift
mkdir -p gpurun_variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --compiler-options -fPIC -Iinclude -Iseqoia_b200/csrc "$@" -shared -o gpurun_variants/libsqoa_b200_$name.so seqoia_b200/csrc/host_api.cu
