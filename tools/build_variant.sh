#!/bin/bash
# tools/build_variant.sh NAME SRC [nvcc flags...]: links a copy of the library whose
# dense_small.cu object is built from SRC with the given flags (A/B timing of kernel
# variants in one GPU call: FBSTAB_B200_LIB=build/variants/NAME.so).
set -e
cd "$(dirname "$0")/.."
name=$1; src=$2; shift 2
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC \
  -Iinclude -Ifbstab_b200/csrc "$@" -c "$src" -o build/variants/$name.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/$name.so \
  build/variants/$name.o build/api.cu.o build/mpc_riccati.cu.o build/mpc_lane.cu.o \
  build/microbench.cu.o build/multi_gpu.cu.o build/closed_loop.cu.o build/problems.cpp.o -ldl
echo built build/variants/$name.so
