set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for c in 2 3a 3b; do timeout 600 python bench.py --config $c --steps 3 --warmup 3 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>gpurun_out/bench_ref.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lane -c 1 -o gpurun_out/lane_servo python tools/prof_mpc.py servo_motor 50 4736 > gpurun_out/ncu_lane.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lane -c 1 -o gpurun_out/lane_dint python tools/prof_mpc.py double_integrator 50 4736 -0.1 >> gpurun_out/ncu_lane.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_small -c 1 -o gpurun_out/dense_small python tools/prof_dense_small.py > gpurun_out/ncu_ds.log 2>&1
ls -la gpurun_out
