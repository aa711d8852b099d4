# compute-sanitizer over every kernel family; the summaries go to profiles/r2_sanitizer.txt
set -x
mkdir -p gpurun_out
OUT=gpurun_out/r2_sanitizer.txt
: > $OUT
run() {  # name, tool, env..., script
  name=$1; tool=$2; shift 2
  echo "==== $name: compute-sanitizer --tool $tool" >> $OUT
  env "$@" timeout 900 compute-sanitizer --tool $tool --print-limit 20 python ${SCRIPT} > gpurun_out/san_tmp.log 2>&1
  echo "exit code $?" >> $OUT
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Error:|=========     at|barrier" gpurun_out/san_tmp.log | head -24 >> $OUT
  grep -v "=========" gpurun_out/san_tmp.log | tail -8 >> $OUT
}
for tool in memcheck racecheck synccheck; do
  SCRIPT=tools/sanitize_mpc.py run "mpc CTA kernel (TMA rings)" $tool X=1
  SCRIPT=tools/sanitize_dense_small.py run "dense small, warp kernel" $tool X=1
done
SCRIPT=tools/sanitize_mpc.py run "mpc lane kernel" memcheck SANITIZE_LANE=1 FBSTAB_MPC_LANE_MIN=256
SCRIPT=tools/sanitize_mpc.py run "mpc lane kernel" racecheck SANITIZE_LANE=1 FBSTAB_MPC_LANE_MIN=256
SCRIPT=tools/sanitize_mpc.py run "mpc lane kernel" synccheck SANITIZE_LANE=1 FBSTAB_MPC_LANE_MIN=256
SCRIPT=tools/sanitize_sparse.py run "sparse lane kernel" memcheck FBSTAB_SPARSE_TEAM=0
SCRIPT=tools/sanitize_sparse.py run "sparse lane kernel" racecheck FBSTAB_SPARSE_TEAM=0
SCRIPT=tools/sanitize_sparse.py run "sparse team kernel" memcheck X=1
SCRIPT=tools/sanitize_sparse.py run "sparse team kernel" racecheck X=1
SCRIPT=tools/sanitize_sparse.py run "sparse team kernel" synccheck X=1
SCRIPT=tools/sanitize_dense_large.py run "dense large kernel" memcheck X=1
SCRIPT=tools/sanitize_dense_large.py run "dense large kernel" racecheck X=1
cat $OUT | tail -150
