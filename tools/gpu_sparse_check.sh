# sparse paths: DRAM traffic capture of config 3a-sparse and compute-sanitizer on both kernels
set -x
mkdir -p gpurun_out/traffic
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'sparse_lane_kernel|sparse_team_kernel' --csv --log-file gpurun_out/traffic/3a-sparse.csv python tools/prof_config.py 3a-sparse > gpurun_out/traffic/3a-sparse.log 2>&1
tail -3 gpurun_out/traffic/3a-sparse.csv | cut -c1-300
for tool in memcheck racecheck synccheck; do
  for team in 1 0; do
    FBSTAB_SPARSE_TEAM=$team timeout 600 compute-sanitizer --tool $tool --print-limit 10 python tools/sanitize_sparse.py > gpurun_out/san_sparse_${tool}_$team.log 2>&1
    echo "== sparse team=$team $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|Error:" gpurun_out/san_sparse_${tool}_$team.log | head -5; grep -v "=========" gpurun_out/san_sparse_${tool}_$team.log | tail -1 | cut -c1-150
  done
done
