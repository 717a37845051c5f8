#!/bin/bash
mkdir -p gpurun_out/r02
for n in 8; do
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n \
    bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02/bench_n${n}_lanes.json 2> gpurun_out/r02/bench_n${n}_lanes.err
  echo "N=$n rc=$?"; tail -c 300 gpurun_out/r02/bench_n${n}_lanes.err
  python - <<P
import json
d=json.loads(open('gpurun_out/r02/bench_n${n}_lanes.json').read().strip().splitlines()[-1])
print('value %.3e e2e %.3e visit %s' % (d['value'], d['e2e']['value'], {k: d['visit'].get(k) for k in ('visits_per_hour','wall_s_max_rank','note','error','lanes_per_gpu')}))
P
done
