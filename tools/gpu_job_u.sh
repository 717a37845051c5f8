#!/bin/bash
# ncu --set full of the boundary update inside a catalogue-field visit CCD, both arithmetic variants
mkdir -p gpurun_out/r02
for a in 1 0; do
  B2_UPDATE_ADDER=$a timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_update_distortions_tiled -s 6 -c 1 \
    -o gpurun_out/r02/prof_update_dense_adder$a -f python tools/visit_kernel_breakdown.py --catalog > gpurun_out/r02/ncu_update_dense_$a.log 2>&1
  echo "adder=$a rc=$?"; tail -3 gpurun_out/r02/ncu_update_dense_$a.log | cut -c1-300
done
