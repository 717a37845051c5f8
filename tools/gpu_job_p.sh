#!/bin/bash
mkdir -p gpurun_out/r02
for n in 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n \
    bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline --no-visit-line > gpurun_out/r02/bench_n${n}_fixed.json 2>/dev/null
  python - <<P
import json
d=json.loads(open('gpurun_out/r02/bench_n${n}_fixed.json').read().strip().splitlines()[-1])
print('N=$n value %.3e e2e %.3e pinned %.3e' % (d['value'], d['e2e']['value'], d['e2e']['pinned_route']['value']))
P
done
