set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sparse.py -x -q -m gpu > gpurun_out/r2_pytest_sp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_sp.log; tail -25 gpurun_out/r2_pytest_sp.log | cut -c1-300
