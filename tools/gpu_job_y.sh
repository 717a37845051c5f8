#!/bin/bash
mkdir -p gpurun_out/r02
SECONDS=0
timeout 1200 python bench.py > gpurun_out/r02/bench_final.json 2> gpurun_out/r02/bench_final.err; echo "bench rc=$? wall ${SECONDS}s"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02/bench_final.json').read().strip().splitlines()[-1])
e=d['e2e']
print('value %.4e e2e %.4e (steps %d, wall %.3f) pinned %.4e visit %.1f' % (d['value'], e['value'], e['steps'], e['wall_s'], e['pinned_route']['value'], d['visit']['visits_per_hour']))
P
