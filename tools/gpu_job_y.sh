#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_visit.py -q -x 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-plugin-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('visit', d['visit'])"
