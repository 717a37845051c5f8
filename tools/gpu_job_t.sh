#!/bin/bash
# A/B of the boundary-update kernels on one box: B2_UPDATE_PULL 0 (32x8 tiles, one slot per thread) / 1 (32x32 tiles, warps pull chunks)
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_sensor.py tests/test_gpu_pool.py tests/test_gpu_edge_cases.py tests/test_gpu_visit.py -q -x 2>&1 | tail -4
for a in 0 1 0 1; do
  echo "== B2_UPDATE_PULL=$a"
  B2_UPDATE_PULL=$a timeout 600 python tools/visit_kernel_breakdown.py --catalog 2>&1 | grep "^R22" | tail -1 | cut -c1-330
  B2_UPDATE_PULL=$a timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-visit-line --no-configs --no-plugin-e2e --kernel-timing 2>&1 | grep -E "per-step kernel" 
done
B2_UPDATE_PULL=1 B2_UPDATE_MODE=0 timeout 600 python tools/visit_kernel_breakdown.py --catalog 2>&1 | grep "^R22" | tail -1 | cut -c1-200
