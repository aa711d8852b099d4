set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "option_caps" > gpurun_out/r2_pytest27.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest27.log; tail -12 gpurun_out/r2_pytest27.log | cut -c1-300
