set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "mpc or lane or closed" > gpurun_out/r2_pytest19.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest19.log; tail -5 gpurun_out/r2_pytest19.log | cut -c1-300
for v in "" lane_s3 lane_nosd; do
  for c in 3a 3b; do
    if [ -z "$v" ]; then timeout 300 python tools/ab_config.py $c 3; else FBSTAB_B200_LIB=build/variants/$v.so timeout 300 python tools/ab_config.py $c 3; fi
  done
done 2>&1 | grep -v "^+" | tee gpurun_out/r2_ab19.txt
FBSTAB_MPC_LANE_CTA_WARPS=3 timeout 300 python tools/ab_config.py 3a 3 2>&1 | grep -v "^+" | tee -a gpurun_out/r2_ab19.txt
FBSTAB_MPC_LANE_CTA_WARPS=2 timeout 300 python tools/ab_config.py 3a 3 2>&1 | grep -v "^+" | tee -a gpurun_out/r2_ab19.txt
