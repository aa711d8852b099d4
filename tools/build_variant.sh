#!/bin/bash
# tools/build_variant.sh NAME SRC [nvcc flags...]: links a copy of the library in which the
# object of SRC (a file of fbstab_b200/csrc, e.g. fbstab_b200/csrc/mpc_lane.cu) is rebuilt
# with the given flags (A/B timing of kernel variants in one GPU call:
# FBSTAB_B200_LIB=build/variants/NAME.so).  The other objects come from build/ (run
# __graft_entry__.build() first).
set -e
cd "$(dirname "$0")/.."
name=$1; src=$2; shift 2
mkdir -p build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC \
  -Iinclude -Ifbstab_b200/csrc "$@" -c "$src" -o build/variants/$name.o
base=$(basename "$src")
objs=""
for o in api.cu dense_small.cu mpc_riccati.cu mpc_lane.cu microbench.cu multi_gpu.cu closed_loop.cu sparse_lane.cu sparse_team.cu problems.cpp sparse_symbolic.cpp; do
  if [ "$o" == "$base" ]; then objs="$objs build/variants/$name.o"; else objs="$objs build/$o.o"; fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/$name.so $objs -ldl
echo built build/variants/$name.so
