#!/bin/bash
# A/B/C of the boundary-update arithmetic on one box (B2_UPDATE_MODE 0 conversions, 1 FP64 adder, 2 double tables + conversion pair)
mkdir -p gpurun_out/r02
for a in 0 1 2 0 1 2; do
  echo "== B2_UPDATE_MODE=$a"
  B2_UPDATE_MODE=$a timeout 600 python tools/visit_kernel_breakdown.py --catalog 2>&1 | grep "^R22" | tail -1 | cut -c1-330
  B2_UPDATE_MODE=$a timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-visit-line --no-configs --no-plugin-e2e --kernel-timing 2>&1 | grep -E "per-step kernel" 
done
timeout 600 python -m pytest tests/test_gpu_sensor.py tests/test_gpu_pool.py -q -x 2>&1 | tail -3
