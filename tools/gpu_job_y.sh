#!/bin/bash
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_gpu_hostpipe.py tests/test_gpu_plugin_pooling.py -q -x 2>&1 | tail -2
for nt in 0 1 0 1; do
  B2_COPY_NT=$nt timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-visit-line --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('NT=$nt e2e %.4e  pinned_route %.4e' % (e['value'], e['pinned_route']['value']))"
done
