#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_plugin_pooling.py -q -x 2>&1 | tail -3
B2_PLUGIN_PROFILE=1 timeout 600 python tools/prof_plugin_e2e.py 2>&1 | tail -2
for k in 1 2; do
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-visit-line --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('e2e %.4e wall %.3f pinned_route %.4e' % (e['value'], e['wall_s'], e['pinned_route']['value']))"
done
