#!/bin/bash
# A/B of the boundary-update arithmetic (FP64-adder rounding vs conversions) on one box
mkdir -p gpurun_out/r02
for a in 1 0 1 0; do
  echo "== B2_UPDATE_ADDER=$a"
  B2_UPDATE_ADDER=$a timeout 600 python tools/visit_kernel_breakdown.py --catalog 2>&1 | grep "^R22" | tail -1 | cut -c1-600
  B2_UPDATE_ADDER=$a timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-visit-line --no-configs --no-plugin-e2e --kernel-timing 2>&1 | grep -E "per-step kernel" 
done
