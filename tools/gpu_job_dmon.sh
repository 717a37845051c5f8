#!/bin/bash
# PCIe counters while eight ranks feed their GPUs through the plugin route (nvidia-smi dmon -s t: rxpci / txpci MB/s)
mkdir -p gpurun_out/r02
nvidia-smi dmon -s t -d 1 -c 45 > gpurun_out/r02/dmon_pcie_n8.txt 2>&1 &
DM=$!
B2_E2E_STEPS=40 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --no-visit-line 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']
print('N=8 e2e %.4e (steps %d) pinned_route %.4e' % (e['value'], e['steps'], e['pinned_route']['value']))"
wait $DM
python - <<'P'
import numpy as np
rows=[l.split() for l in open('gpurun_out/r02/dmon_pcie_n8.txt') if l.strip() and not l.startswith('#')]
a=np.array([[int(r[0]), float(r[1]), float(r[2])] for r in rows if len(r)>=3])
for g in range(8):
    m=a[a[:,0]==g]
    print('gpu %d rxpci max %.0f MB/s  p90 %.0f  txpci max %.0f' % (g, m[:,1].max(), np.percentile(m[:,1],90), m[:,2].max()))
tot=[a[(a[:,0]>=0)][i::8,1].sum() for i in range(0)]
P
