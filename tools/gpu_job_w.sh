#!/bin/bash
mkdir -p gpurun_out/r02
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_run.py > gpurun_out/r02/sanitizer2_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r02/sanitizer2_$tool.log | tail -1)"; grep -E "^stamps" gpurun_out/r02/sanitizer2_$tool.log | tail -2
done
# ncu of the cluster launch of the stamps kernel on the classic catalogue
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stamp_jobs -c 2 \
    -o gpurun_out/r02/prof_stamps_cluster -f python tools/classic_bench.py 1998 5e7 > gpurun_out/r02/ncu_stamps_cluster.log 2>&1
echo "ncu rc=$?"; grep "^build " gpurun_out/r02/ncu_stamps_cluster.log | tail -1 | cut -c1-200
