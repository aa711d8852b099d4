set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "mpc or lane or closed" > gpurun_out/r2_pytest21.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest21.log; tail -5 gpurun_out/r2_pytest21.log | cut -c1-300
for c in 3a 3b; do timeout 300 python tools/ab_config.py $c 3; done 2>&1 | grep -v "^+" | tee gpurun_out/r2_ab21.txt
timeout 600 python tests/closed_loop_bench.py > gpurun_out/r2_closed_loop2.txt 2>&1; tail -3 gpurun_out/r2_closed_loop2.txt | cut -c1-400
