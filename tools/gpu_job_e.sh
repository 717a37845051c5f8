#!/bin/bash
mkdir -p gpurun_out/r02
timeout 300 python -m pytest tests/test_gpu_stamps.py tests/test_gpu_classic.py -x -q 2>&1 | tail -30
echo "--- classic bench"; timeout 600 python tools/classic_bench.py > gpurun_out/r02/classic_bench.log 2>&1; tail -5 gpurun_out/r02/classic_bench.log
B2_TIMING=1 timeout 300 python - <<'PY' 2>&1 | tail -5
import sys, runpy, json
sys.argv=["tools/classic_bench.py"]
try:
    runpy.run_path("tools/classic_bench.py", run_name="__main__")
except SystemExit: pass
from imsim_b200._lib import timing_report
print(json.dumps({k:[v[0], round(v[1],2)] for k,v in sorted(timing_report().items(), key=lambda kv:-kv[1][1])[:12]}))
PY
