set -x
mkdir -p gpurun_out
FBSTAB_DENSE_SMALL_TEAM=2 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dense" > gpurun_out/r2_pytest7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest7.log; tail -25 gpurun_out/r2_pytest7.log | cut -c1-220
for t in 1 2 2; do FBSTAB_DENSE_SMALL_TEAM=$t timeout 300 python tools/ab_dense_small.py 65536 4 2>&1 | tail -1 | tee -a gpurun_out/r2_ab7.txt; done
