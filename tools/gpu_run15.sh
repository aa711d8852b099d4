set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mpc_lane_kernel -s 2 -c 1 -o gpurun_out/r2_lane_shared python tools/prof_config.py 3a > gpurun_out/r2_ncu_lane.log 2>&1
tail -3 gpurun_out/r2_ncu_lane.log
