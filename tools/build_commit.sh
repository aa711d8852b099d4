#!/bin/bash
# tools/build_commit.sh COMMIT NAME: builds the engine library of a past commit into
# build/variants/NAME.so (A/B timing and bit-identity checks against the working tree:
# FBSTAB_B200_LIB=build/variants/NAME.so).
set -e
cd "$(dirname "$0")/.."
commit=$1; name=$2
d=build/variants/src_$name
rm -rf $d; mkdir -p $d/csrc $d/include
git archive $commit fbstab_b200/csrc include | tar -x -C $d
objs=""
for f in $d/fbstab_b200/csrc/*.cu $d/fbstab_b200/csrc/problems.cpp; do
  o=$d/$(basename $f).o
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC \
    -I$d/include -I$d/fbstab_b200/csrc -c $f -o $o &
  objs="$objs $o"
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/$name.so $objs -ldl
echo built build/variants/$name.so
