#!/bin/bash
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02/pytest_gpu_4.log 2>&1; echo "suite rc=$?"; tail -6 gpurun_out/r02/pytest_gpu_4.log
echo "--- fused vs unfused"; timeout 200 python tools/diag_fused.py 2>&1 | tail -6
echo "--- classic bench"; timeout 600 python tools/classic_bench.py > gpurun_out/r02/classic_bench.log 2>&1; tail -6 gpurun_out/r02/classic_bench.log
