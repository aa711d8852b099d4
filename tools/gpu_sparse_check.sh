set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sparse.py -x -q -m gpu > gpurun_out/r2_pytest_sp.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_sp.log; tail -5 gpurun_out/r2_pytest_sp.log | cut -c1-300
timeout 600 python tools/time_sparse.py servo_motor 50 16384 2>gpurun_out/sp.err | cut -c1-560; tail -3 gpurun_out/sp.err
