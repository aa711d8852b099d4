set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
for c in 2 3a 3b 4a 4b 5; do timeout 600 python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/final_$c.json 2> gpurun_out/final_$c.err; done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_ref.json 2>/dev/null
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_cfg2.csv python bench.py --steps 2 --warmup 1 > gpurun_out/launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_small -s 1 -c 1 -o gpurun_out/dense_small8 python tools/prof_dense_small.py 16384 > gpurun_out/ncu_ds8.log 2>&1
ls -la gpurun_out | tail -20
