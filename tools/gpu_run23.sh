set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sparse.py -x -q -m gpu > gpurun_out/r2_pytest23.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest23.log; tail -4 gpurun_out/r2_pytest23.log | cut -c1-300
timeout 600 python tools/time_sparse.py servo_motor 50 16384 > gpurun_out/r2_sparse_servo.json 2> gpurun_out/r2_sparse_servo.err; tail -c 900 gpurun_out/r2_sparse_servo.json; tail -5 gpurun_out/r2_sparse_servo.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sparse_lane_kernel -s 1 -c 1 -o gpurun_out/r2_sparse_lane python tools/time_sparse.py servo_motor 50 16384 > gpurun_out/r2_ncu_sparse.log 2>&1
