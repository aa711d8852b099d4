set -x
mkdir -p gpurun_out
FBSTAB_B200_LIB=build/variants/ds_single.so timeout 300 python tools/ab_config.py 2 4 2>&1 | grep -v "^+" | tee gpurun_out/r2_ab26.txt
timeout 300 python tools/ab_config.py 2 4 2>&1 | grep -v "^+" | tee -a gpurun_out/r2_ab26.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "dense" > gpurun_out/r2_pytest26.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest26.log; tail -4 gpurun_out/r2_pytest26.log | cut -c1-300
