#!/bin/bash
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_stamps.py tests/test_gpu_classic.py -q -x 2>&1 | tail -15 > gpurun_out/r02/pytest_stamps_cluster.log
echo "stamps rc=$?"; tail -8 gpurun_out/r02/pytest_stamps_cluster.log
for cs in 1 8 4; do
  echo "== cluster $cs"
  B2_STAMP_CLUSTER=$cs timeout 600 python tools/classic_bench.py 1998 5e7 2>&1 | grep -v "^{" | grep "^build " | tail -2
done
echo "== profile cluster 8"
B2_CLASSIC_PROFILE=1 B2_STAMP_PROFILE=1 timeout 600 python tools/classic_bench.py 1998 5e7 2>&1 | grep -E "cycles|classic_bench" | tail -4
echo "== heavy threshold sweep (cluster 8)"
for h in 5e4 1e5 4e5 1e6; do
  echo "heavy $h"; B2_STAMP_HEAVY=$h timeout 600 python tools/classic_bench.py 1998 5e7 2>&1 | grep "^build " | tail -1
done
