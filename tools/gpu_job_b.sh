#!/bin/bash
# round-2 GPU job B: stamps kernel tests first (under a timeout: new persistent kernel), then the whole gpu suite
mkdir -p gpurun_out/r02
timeout 300 python -m pytest tests/test_gpu_stamps.py -x -q > gpurun_out/r02/pytest_stamps.log 2>&1; echo "stamps rc=$?"; tail -25 gpurun_out/r02/pytest_stamps.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02/pytest_gpu_3.log 2>&1; echo "suite rc=$?"; tail -8 gpurun_out/r02/pytest_gpu_3.log
