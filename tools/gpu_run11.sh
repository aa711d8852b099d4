set -x
mkdir -p gpurun_out
for c in 3a 3b 4b 4a40 5; do
  FBSTAB_B200_LIB=$PWD/build/variants/prev.so timeout 600 python tools/ab_config.py $c 2>&1 | tail -1 | tee -a gpurun_out/r2_ab11.txt
  timeout 600 python tools/ab_config.py $c 2>&1 | tail -1 | tee -a gpurun_out/r2_ab11.txt
done
