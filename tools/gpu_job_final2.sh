#!/bin/bash
# evidence of the final build: launch list of the bench step, full capture of the boundary update (mode 2) in it
mkdir -p gpurun_out/r02
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02/launches_bench_final.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-visit-line --no-configs --no-plugin-e2e > gpurun_out/r02/ncu_launches_final.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r02/launches_bench_final.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_update_distortions_tiled -s 2 -c 1 \
  -o gpurun_out/r02/prof_update_tiled_final -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-visit-line --no-configs --no-plugin-e2e > gpurun_out/r02/ncu_update_final.log 2>&1
echo "update capture rc=$?"
