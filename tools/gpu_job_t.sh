#!/bin/bash
# A/B: horizontal and vertical slots of a tile in one block (0) or two (1)
for a in 0 1 0 1; do
  echo "== B2_UPDATE_SPLIT=$a"
  B2_UPDATE_SPLIT=$a timeout 600 python tools/visit_kernel_breakdown.py --catalog 2>&1 | grep "^R22" | tail -1 | cut -c1-120
  B2_UPDATE_SPLIT=$a timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-visit-line --no-configs --no-plugin-e2e --kernel-timing 2>&1 | grep -E "per-step kernel" 
done
B2_UPDATE_SPLIT=1 timeout 900 python -m pytest tests/test_gpu_sensor.py tests/test_gpu_pool.py -q -x 2>&1 | tail -2
