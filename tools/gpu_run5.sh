set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest5.log; tail -5 gpurun_out/r2_pytest5.log
( time timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err ) 2>&1 | tail -3; tail -12 gpurun_out/r2_bench5.err; wc -c gpurun_out/r2_bench5.json
timeout 600 python tests/closed_loop_bench.py > gpurun_out/r2_closed_loop.txt 2>&1; tail -3 gpurun_out/r2_closed_loop.txt
