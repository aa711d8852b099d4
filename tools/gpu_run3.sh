set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest3.log; tail -5 gpurun_out/r2_pytest3.log
for v in v2 v3 v3; do FBSTAB_B200_LIB=$PWD/build/variants/$v.so timeout 300 python tools/ab_dense_small.py 65536 4 2>&1 | tail -1 | tee -a gpurun_out/r2_ab3.txt; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_small -s 1 -c 1 -o gpurun_out/r2_dense_small_v3 python tools/prof_dense_small.py 16384 > gpurun_out/r2_ncu_v3.log 2>&1
