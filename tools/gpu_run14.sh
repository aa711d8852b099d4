# lane kernel with the L1-bypassing workspace: prefetch distance A/B, ncu source-level capture
set -x
mkdir -p gpurun_out
for v in "" lane_pd2 lane_pd3; do
  for c in 3a 3b; do
    if [ -z "$v" ]; then timeout 300 python tools/ab_config.py $c 3; else FBSTAB_B200_LIB=build/variants/$v.so timeout 300 python tools/ab_config.py $c 3; fi
  done
done 2>&1 | grep -v "^+" | tee gpurun_out/r2_ab14.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mpc_lane_kernel<4, 1, 4, 1>' -s 1 -c 1 -o gpurun_out/r2_lane_shared python tools/prof_config.py 3a > gpurun_out/r2_ncu_lane.log 2>&1
tail -3 gpurun_out/r2_ncu_lane.log
