#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_plugin_pooling.py tests/test_plugin_wiring.py tests/test_gpu_pool.py -x -q 2>&1 | tail -30
