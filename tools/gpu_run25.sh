set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sparse.py tests/test_cpp_facade.py -x -q -m gpu > gpurun_out/r2_pytest25.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest25.log; tail -4 gpurun_out/r2_pytest25.log | cut -c1-300
timeout 900 python bench.py --config 3a-sparse --per-config none --steps 2 --warmup 3 > gpurun_out/r2_bench_sparse.json 2> gpurun_out/r2_bench_sparse.err; tail -3 gpurun_out/r2_bench_sparse.err; head -c 3000 gpurun_out/r2_bench_sparse.json
