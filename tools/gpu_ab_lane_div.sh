# A/B of the MPC lane kernel's factor diagonals (FBSTAB_LANE_EXACT_DIV): 0 = reciprocal by
# rsqrt, multiply; 1 = IEEE sqrt, divide (the reference's operation sequence); 2 = 1 / sqrt
# correctly rounded, multiply; 3 = rsqrt + one Newton step, multiply -- throughput and
# trajectory agreement with the oracle on cfg 3a / 3b.  Usage: bash tools/gpu_ab_lane_div.sh tree lane_div2 ...
mkdir -p gpurun_out
for v in "$@"; do
  unset FBSTAB_B200_LIB
  if [ $v != tree ]; then export FBSTAB_B200_LIB=build/variants/$v.so; fi
  for c in 3a 3b; do
    python bench.py --config $c --per-config none --steps 3 --warmup 3 > gpurun_out/ab_lane_${v}_$c.json 2> gpurun_out/ab_lane_${v}_$c.err
  done
  python - <<P
import json
for c in ("3a","3b"):
    d=json.load(open("gpurun_out/ab_lane_${v}_%s.json"%c)); cb=d["cpu_baseline"]
    print("$v", c, "%.4g solves/s"%d["value"], {k:cb[k] for k in cb if "traj" in k or "flags" in k or "diff" in k or "newton" in k})
P
done
