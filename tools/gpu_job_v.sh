#!/bin/bash
mkdir -p gpurun_out/r02
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02/pytest_gpu_6.log
echo "suite rc=$?"; tail -12 gpurun_out/r02/pytest_gpu_6.log
