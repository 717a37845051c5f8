#!/bin/bash
mkdir -p gpurun_out/r02
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02/pytest_gpu_5.log
echo "suite rc=$?"; tail -6 gpurun_out/r02/pytest_gpu_5.log
timeout 600 python tools/classic_bench.py 1998 5e7 2>&1 | grep "^build " | tail -2
B2_CLASSIC_PROFILE=1 B2_STAMP_PROFILE=1 timeout 600 python tools/classic_bench.py 1998 5e7 2>&1 | grep -E "cycles|classic_bench" | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --kernel-timing > gpurun_out/r02/bench_round_fp64.json 2> gpurun_out/r02/bench_round_fp64.err
tail -c 400 gpurun_out/r02/bench_round_fp64.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02/bench_round_fp64.json').read().strip().splitlines()[-1])
print('value %.4e ms %.3f'%(d['value'], d['ms_per_step']), d.get('breakdown_ms'), d.get('kernel_timing'))
print('visit', d.get('visit'))
print('e2e', d['e2e']['value'])
print('configs', d.get('configs'))
P
