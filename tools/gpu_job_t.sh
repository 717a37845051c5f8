#!/bin/bash
# A/B of the boundary-update table placement: B2_UPDATE_MODE 2 (double tables through L1) / 3 (padded, in shared memory)
for a in 2 3 2 3; do
  echo "== B2_UPDATE_MODE=$a"
  B2_UPDATE_MODE=$a timeout 600 python tools/visit_kernel_breakdown.py --catalog 2>&1 | grep "^R22" | tail -1 | cut -c1-200
  B2_UPDATE_MODE=$a timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-visit-line --no-configs --no-plugin-e2e --kernel-timing 2>&1 | grep -E "per-step kernel" 
done
B2_UPDATE_MODE=3 timeout 900 python -m pytest tests/test_gpu_sensor.py tests/test_gpu_pool.py tests/test_gpu_edge_cases.py -q -x 2>&1 | tail -2
