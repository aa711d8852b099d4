set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sparse.py -x -q -m gpu > gpurun_out/r2_pytest24.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest24.log; tail -4 gpurun_out/r2_pytest24.log | cut -c1-300
for L in 32 16 8; do FBSTAB_SPARSE_LANES=$L timeout 600 python tools/time_sparse.py servo_motor 50 16384 2>/dev/null | cut -c1-520; done | tee gpurun_out/r2_sparse_lanes.txt
