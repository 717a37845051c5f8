#!/bin/bash
# what the driver runs at round end: gpu tests, smoke, the default bench line
mkdir -p gpurun_out/r02
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02/pytest_gpu_final.log; tail -3 gpurun_out/r02/pytest_gpu_final.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
SECONDS=0
timeout 1200 python bench.py > gpurun_out/r02/bench_final.json 2> gpurun_out/r02/bench_final.err; echo "bench rc=$? wall ${SECONDS}s"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02/bench_final.json').read().strip().splitlines()[-1])
print('value %.4e ms %.3f e2e %.4e launches %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']))
print('roofline', d['roofline']['frac'], d['roofline_fp64']['frac'], 'breakdown', d['breakdown_ms'])
print('visit', d['visit']['visits_per_hour'], 'clocks', d['clocks'])
print({k:'%.3e'%v['value'] for k,v in d['configs'].items()})
P
