#!/bin/bash
mkdir -p gpurun_out/r02
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:^k_accumulate$' -s 1 -c 1 -o gpurun_out/r02/prof_accumulate -f python bench.py --unfused --steps 2 --warmup 3 --no-visit-line --no-cpu-baseline --no-configs --no-plugin-e2e > gpurun_out/r02/ncu_accumulate.log 2>&1; tail -2 gpurun_out/r02/ncu_accumulate.log | cut -c1-200; ls -la gpurun_out/r02/prof_accumulate.ncu-rep
