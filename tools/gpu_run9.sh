set -x
mkdir -p gpurun_out
for v in v5 v6 v7 v8 v5; do FBSTAB_B200_LIB=$PWD/build/variants/$v.so timeout 300 python tools/ab_dense_small.py 65536 4 2>&1 | tail -1 | tee -a gpurun_out/r2_ab9.txt; done
