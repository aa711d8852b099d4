# lane-kernel A/B: workspace cache policy and common stage data in shared memory; fresh ncu
# source-level capture of the shared-data lane kernel on the full cfg 3a batch
set -x
mkdir -p gpurun_out
for v in "" lane_cg lane_cs lane_smem lane_cg_smem; do
  for c in 3a 3b; do
    if [ -z "$v" ]; then timeout 300 python tools/ab_config.py $c 3; else FBSTAB_B200_LIB=build/variants/$v.so timeout 300 python tools/ab_config.py $c 3; fi
  done
done 2>&1 | grep -v "^+" | tee gpurun_out/r2_ab13.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mpc_lane_kernel -s 1 -c 1 -o gpurun_out/r2_lane_shared python tools/prof_config.py 3a > gpurun_out/r2_ncu_lane.log 2>&1
tail -3 gpurun_out/r2_ncu_lane.log
