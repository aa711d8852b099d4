set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "mpc or lane or closed" > gpurun_out/r2_pytest20.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest20.log; tail -5 gpurun_out/r2_pytest20.log | cut -c1-300
for c in 3a 3b; do timeout 300 python tools/ab_config.py $c 3; done 2>&1 | grep -v "^+" | tee gpurun_out/r2_ab20.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mpc_lane_kernel -s 2 -c 1 -o gpurun_out/r2_lane_ring2 python tools/prof_config.py 3a > gpurun_out/r2_ncu_lane.log 2>&1
