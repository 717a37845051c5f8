#!/bin/bash
B2_PIPE_PROFILE=1 python bench.py --no-visit-line --no-cpu-baseline --steps 3 2>&1 >/dev/null | grep photons_upload | tail -4
B2_PIPE_PROFILE=1 B2_HOST_THREADS=4 python bench.py --no-visit-line --no-cpu-baseline --steps 3 2>&1 >/dev/null | grep photons_upload | tail -2
