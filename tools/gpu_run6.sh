set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest6.log; tail -30 gpurun_out/r2_pytest6.log | cut -c1-250
