#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_stamps.py -x -q 2>&1 | tail -3
B2_STAMP_PROFILE=1 timeout 600 python tools/classic_bench.py 2>&1 | grep -v per_object | tail -4 | cut -c1-400
