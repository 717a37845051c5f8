#!/bin/bash
# round-2 GPU job J: ncu evidence -- launch list of the bench step, full captures of k_accumulate (unfused bench) and k_stamp_jobs
mkdir -p gpurun_out/r02
F="--no-visit-line --no-cpu-baseline --no-configs --no-plugin-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02/launches_bench.csv python bench.py --steps 2 --warmup 3 $F > gpurun_out/r02/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_accumulate\$ -s 3 -c 1 -o gpurun_out/r02/prof_accumulate -f python bench.py --unfused --steps 2 --warmup 3 $F > gpurun_out/r02/ncu_accumulate.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_update_distortions_tiled -s 3 -c 1 -o gpurun_out/r02/prof_update_tiled -f python bench.py --steps 2 --warmup 3 $F > gpurun_out/r02/ncu_update.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stamp_jobs -s 1 -c 1 -o gpurun_out/r02/prof_stamps -f python tools/classic_bench.py 1998 5e7 > gpurun_out/r02/ncu_stamps.log 2>&1
ls -la gpurun_out/r02/*.ncu-rep; tail -2 gpurun_out/r02/ncu_stamps.log
