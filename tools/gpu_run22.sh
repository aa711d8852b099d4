set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sparse.py -x -q -m gpu > gpurun_out/r2_pytest22.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest22.log; tail -25 gpurun_out/r2_pytest22.log | cut -c1-300
timeout 600 python tools/time_sparse.py servo_motor 50 16384 > gpurun_out/r2_sparse_servo.json 2> gpurun_out/r2_sparse_servo.err; tail -c 1500 gpurun_out/r2_sparse_servo.json; tail -5 gpurun_out/r2_sparse_servo.err
