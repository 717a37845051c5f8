#!/bin/bash
mkdir -p gpurun_out/r02
for nt in 1 0; do
  B2_COPY_NT=$nt timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2952$nt \
    bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --no-visit-line 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('N=8 NT=$nt value %.4e e2e %.4e  pinned_route %.4e' % (d['value'], e['value'], e['pinned_route']['value']))"
done
