set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "mpc or lane or closed" > gpurun_out/r2_pytest17.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest17.log; tail -5 gpurun_out/r2_pytest17.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mpc_lane_kernel -s 2 -c 1 -o gpurun_out/r2_lane_ring python tools/prof_config.py 3a > gpurun_out/r2_ncu_lane.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mpc_lane_kernel -s 2 -c 1 -o gpurun_out/r2_lane_ring_3b python tools/prof_config.py 3b > gpurun_out/r2_ncu_lane3b.log 2>&1
