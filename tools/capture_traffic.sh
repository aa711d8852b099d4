#!/bin/bash
# DRAM traffic of the dominant kernel of every BASELINE config (the `roofline.traffic`
# of bench.py): ncu dram__bytes_read.sum + dram__bytes_write.sum of the second launch,
# per instance, written to profiles/traffic.json together with a hash of the kernel
# sources -- bench.py prints the figure only while that hash matches the build it runs.
# Usage (GPU box): bash tools/capture_traffic.sh
set -x
mkdir -p gpurun_out/traffic
for c in 2 3a 3b 4a 4a40 4b 5 3a-sparse; do
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    --clock-control none -k regex:'dense_small_kernel|mpc_lane_kernel|mpc_riccati_kernel|dense_large_kernel|sparse_lane_kernel|sparse_team_kernel' \
    --csv --log-file gpurun_out/traffic/$c.csv python tools/prof_config.py $c > gpurun_out/traffic/$c.log 2>&1
done
python tools/traffic_json.py gpurun_out/traffic profiles/traffic.json
cat profiles/traffic.json
cp profiles/traffic.json gpurun_out/traffic.json
