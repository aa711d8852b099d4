# A/B of the MPC CTA kernel's Cholesky diagonal (FBSTAB_CTA_DIAG: 0 = rsqrt, 1 = 1 / sqrt
# correctly rounded): throughput and trajectory agreement with the oracle on the servo
# problem forced onto the CTA kernel, cfg 4a40 and cfg 4b.  Usage: bash tools/gpu_ab_cta_diag.sh tree cta_diag1
mkdir -p gpurun_out
for v in "$@"; do
  unset FBSTAB_B200_LIB
  if [ $v != tree ]; then export FBSTAB_B200_LIB=build/variants/$v.so; fi
  FBSTAB_MPC_LANE=0 python bench.py --config 3a --per-config none --steps 2 --warmup 3 > gpurun_out/ab_cta_${v}_3a.json 2> gpurun_out/ab_cta_${v}_3a.err
  for c in 4a40 4b; do
    python bench.py --config $c --per-config none --steps 2 --warmup 3 > gpurun_out/ab_cta_${v}_$c.json 2> gpurun_out/ab_cta_${v}_$c.err
  done
  python - <<P
import json
for c in ("3a","4a40","4b"):
    d=json.load(open("gpurun_out/ab_cta_${v}_%s.json"%c)); cb=d["cpu_baseline"]
    print("$v", c, d["config"].get("kernel_path", ""), "%.4g solves/s"%d["value"], {k:cb[k] for k in cb if "traj" in k or "flags" in k or "diff" in k or "newton" in k})
P
done
