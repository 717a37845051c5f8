#!/bin/bash
# the default bench line as the driver runs it
mkdir -p gpurun_out/r02
SECONDS=0
timeout 1200 python bench.py > gpurun_out/r02/bench_final.json 2> gpurun_out/r02/bench_final.err; echo "bench rc=$? wall ${SECONDS}s"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02/bench_final.json').read().strip().splitlines()[-1])
print('value %.4e ms %.3f e2e %.4e launches %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']))
print('roofline', d['roofline']['frac'], d['roofline_fp64']['frac'], 'cpu', d['cpu_baseline'])
print('visit', d['visit']['visits_per_hour'], 'clocks', d['clocks'])
print({k:'%.3e'%v['value'] for k,v in d['configs'].items()})
P
