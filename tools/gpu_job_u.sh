#!/bin/bash
# stage 1: blocks per SM held by the register allocation (B2_STAGE1_OCC 4 / 5 / 6 / 8), per-kernel time in a catalogue-field CCD
for o in 4 5 6 8 4 6; do
  echo "== B2_STAGE1_OCC=$o"
  B2_STAGE1_OCC=$o timeout 600 python tools/visit_kernel_breakdown.py --catalog 2>&1 | grep "^R22" | tail -1 | grep -o "'k_stage1_photons': ([0-9]*, [0-9.]*)"
done
B2_STAGE1_OCC=6 timeout 600 python -m pytest tests/test_gpu_stage1.py -q -x 2>&1 | tail -1
