set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest4.log; tail -5 gpurun_out/r2_pytest4.log
for v in v3 v4 v4; do FBSTAB_B200_LIB=$PWD/build/variants/$v.so timeout 300 python tools/ab_dense_small.py 65536 4 2>&1 | tail -1 | tee -a gpurun_out/r2_ab4.txt; done
timeout 600 python tests/closed_loop_bench.py > gpurun_out/r2_closed_loop.txt 2>&1; tail -3 gpurun_out/r2_closed_loop.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > gpurun_out/r2_multi_gpu_check.txt 2>&1; tail -8 gpurun_out/r2_multi_gpu_check.txt
