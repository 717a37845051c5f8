#!/bin/bash
mkdir -p gpurun_out/r02
timeout 300 python -m pytest tests/test_gpu_stamps.py -x -q 2>&1 | tail -5
echo "--- fused vs unfused, compiled XyToV"; timeout 200 python tools/diag_fused.py 2>&1 | tail -8
echo "--- fused vs unfused, exact XyToV"; B2_XYTOV_EXACT=1 timeout 200 python tools/diag_fused.py 2>&1 | tail -8
