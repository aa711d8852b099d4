set -x
mkdir -p gpurun_out
FBSTAB_DENSE_SMALL_TEAM=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_small2 -s 1 -c 1 -o gpurun_out/r2_dense_small2_a python tools/prof_dense_small.py 16384 > gpurun_out/r2_ncu_ds2a.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
