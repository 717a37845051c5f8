#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_stamps.py tests/test_gpu_classic.py -q -x 2>&1 | tail -2
timeout 600 python tools/classic_bench.py 1998 5e7 2>&1 | grep "^build " | tail -2 | cut -c1-160
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-visit-line 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:'%.3e'%v['value'] for k,v in d['configs'].items()}, 'e2e %.3e'%d['e2e']['value'])"
