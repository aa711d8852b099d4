set -x
mkdir -p gpurun_out
for co in "" 100 50; do
for c in 3a 3b; do FBSTAB_MPC_LANE_CARVEOUT=$co timeout 300 python tools/ab_config.py $c 3; done
done 2>&1 | grep -v "^+" | tee gpurun_out/r2_ab18.txt
