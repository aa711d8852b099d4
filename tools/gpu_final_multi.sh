# Multi-GPU measurement (gpurun --gpus N): the NCCL gather check and the bench line at N ranks.
set -x
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > gpurun_out/r2_multi_gpu_check_$N.txt 2>&1; tail -6 gpurun_out/r2_multi_gpu_check_$N.txt
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 5 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err ) 2>&1 | tail -3
grep "\[bench\]" gpurun_out/r2_bench_${N}gpu.err | tail -9
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 5 --scaling strong --per-config none > gpurun_out/r2_bench_${N}gpu_strong.json 2> gpurun_out/r2_bench_${N}gpu_strong.err ) 2>&1 | tail -3
wc -c gpurun_out/r2_bench_${N}gpu*.json
