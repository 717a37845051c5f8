#!/bin/bash
mkdir -p gpurun_out/r02
for tool in memcheck racecheck initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_run.py > gpurun_out/r02/sanitizer_$tool.log 2>&1
  echo "== $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/r02/sanitizer_$tool.log | tail -1)"; grep -E "^stamps|upload / host" gpurun_out/r02/sanitizer_$tool.log | tail -2
done
