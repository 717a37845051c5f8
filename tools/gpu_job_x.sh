#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_optics.py tests/test_gpu_pool.py tests/test_spike_statistics.py -q -x -m gpu 2>&1 | tail -2
for r in 1 0 1 0; do
  B2_TRACE_ROLLED=$r timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-visit-line --no-configs --no-plugin-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('rolled=$r value %.4e ms %.3f k_pool_step %.3f frac %.3f' % (d['value'], d['ms_per_step'], d['breakdown_ms']['k_pool_step'], d['roofline_fp64']['frac']))"
done
