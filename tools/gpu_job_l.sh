#!/bin/bash
mkdir -p gpurun_out/r02
python -m pytest tests/test_gpu_sensor.py tests/test_gpu_pool.py tests/test_gpu_visit.py tests/test_gpu_stamps.py -x -q 2>&1 | tail -4
python bench.py --no-visit-line --no-cpu-baseline --no-configs --no-plugin-e2e --steps 8 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value', d['value'], d['ms_per_step'], d['breakdown_ms'])"
python tools/visit_kernel_breakdown.py --catalog 2>&1 | tail -1 | cut -c1-500
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:^k_accumulate$' -s 3 -c 1 -o gpurun_out/r02/prof_accumulate -f python bench.py --unfused --steps 2 --warmup 3 --no-visit-line --no-cpu-baseline --no-configs --no-plugin-e2e > gpurun_out/r02/ncu_accumulate.log 2>&1; tail -1 gpurun_out/r02/ncu_accumulate.log
