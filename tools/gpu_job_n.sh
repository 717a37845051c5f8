#!/bin/bash
mkdir -p gpurun_out/r02
python -m pytest tests/test_gpu_plugin_pooling.py -x -q 2>&1 | tail -15
python bench.py --no-visit-line --no-cpu-baseline --no-configs --steps 6 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value', d['value'], 'e2e', d['e2e'].get('value'), d['e2e'].get('wall_s'), d['e2e'].get('plugin_route'))"
