# lane kernel with the TMA stage-block ring: parity tests, A/B (bit-identity hash), sanitizer
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "mpc or lane or closed" > gpurun_out/r2_pytest16.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest16.log; tail -5 gpurun_out/r2_pytest16.log | cut -c1-300
for c in 3a 3b; do timeout 300 python tools/ab_config.py $c 3; done 2>&1 | grep -v "^+" | tee gpurun_out/r2_ab16.txt
for tool in memcheck racecheck synccheck; do
  SANITIZE_LANE=1 FBSTAB_MPC_LANE_MIN=256 timeout 600 compute-sanitizer --tool $tool --print-limit 10 python tools/sanitize_mpc.py > gpurun_out/san_lane_$tool.log 2>&1
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Error:" gpurun_out/san_lane_$tool.log | head -8
done
