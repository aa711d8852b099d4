set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest10.log; tail -5 gpurun_out/r2_pytest10.log | cut -c1-250
for c in 4a40 4b 4a; do timeout 600 python bench.py --config $c --per-config none --no-cpu --steps 2 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['baseline_config'], d['value'], d['results']['path'][:40])" | tee -a gpurun_out/r2_mpc_inplace.txt; done
FBSTAB_MPC_LANE=0 timeout 600 python bench.py --config 3a --per-config none --no-cpu --steps 2 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('3a-cta', d['value'], d['results']['path'][:40])" | tee -a gpurun_out/r2_mpc_inplace.txt
bash tools/gpu_sanitize.sh > gpurun_out/r2_sanitize_run.log 2>&1; tail -80 gpurun_out/r2_sanitizer.txt
