set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest12.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest12.log; tail -5 gpurun_out/r2_pytest12.log | cut -c1-250
timeout 600 python tests/closed_loop_bench.py > gpurun_out/r2_closed_loop.txt 2>&1; tail -3 gpurun_out/r2_closed_loop.txt | cut -c1-600
