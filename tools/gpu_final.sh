# The end-of-round measurement sequence on ONE GPU (GPU tests, smoke, the full bench line,
# the reference arm, the ncu launch list of the bench command and the full capture of the
# headline kernel, DRAM traffic of every config).  Multi-GPU: tools/gpu_final_multi.sh.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_final.log
tail -3 gpurun_out/r2_pytest_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
bash tools/capture_traffic.sh > gpurun_out/r2_traffic.log 2>&1; tail -5 gpurun_out/r2_traffic.log
( time timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err ) 2>&1 | tail -3
tail -9 gpurun_out/r2_bench_final.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_ref.json 2>/dev/null
timeout 600 python tests/closed_loop_bench.py > gpurun_out/r2_closed_loop.txt 2>&1; tail -2 gpurun_out/r2_closed_loop.txt | cut -c1-400
for B in 16384 65536; do timeout 600 python tests/sparse_bench.py servo_motor 50 $B 2>/dev/null; done > gpurun_out/r2_sparse_lane.txt; cut -c1-420 gpurun_out/r2_sparse_lane.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --per-config 3a,4b --no-cpu > gpurun_out/r2_launches_bench.log 2>&1
# (the full ncu capture of the headline kernel -- profiles/r2_dense_small_final_ncu.txt -- is of the same kernel source:
#  timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_small -s 1 -c 1 -o gpurun_out/r2_dense_small_final python tools/prof_dense_small.py 16384)
ls -la gpurun_out | tail -12
