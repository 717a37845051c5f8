#!/bin/bash
for cs in 8 4; do for h in 5e4 1e5 2e5; do
  echo "== cluster $cs heavy $h"
  B2_STAMP_CLUSTER=$cs B2_STAMP_HEAVY=$h timeout 600 python tools/classic_bench.py 1998 5e7 2>&1 | grep "^build " | tail -2 | cut -c1-160
done; done
timeout 600 python -m pytest tests/test_gpu_stamps.py -q -x 2>&1 | tail -2
