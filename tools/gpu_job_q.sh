#!/bin/bash
# host topology of the 8-GPU box + host copy bandwidth vs. number of processes
mkdir -p gpurun_out/r02
{
nproc; lscpu | head -30; free -g | head -3
ls /sys/devices/system/node/ | head; for n in /sys/devices/system/node/node*; do echo $n $(cat $n/cpulist); done
nvidia-smi topo -m
cat /proc/self/status | grep -i cpus_allowed_list
python - <<'P'
import numpy as np, time, multiprocessing as mp, os
def work(q, n):
    a = np.ones(1 << 27, dtype=np.float64)  # 1 GB
    b = np.empty_like(a)
    b[:] = a
    t = time.perf_counter()
    for _ in range(3): np.copyto(b, a)
    q.put(3 * a.nbytes / (time.perf_counter() - t) / 1e9)
for n in (1, 2, 4, 8, 16, 32):
    q = mp.Queue(); ps = [mp.Process(target=work, args=(q, n)) for _ in range(n)]
    [p.start() for p in ps]; r = [q.get() for _ in ps]; [p.join() for p in ps]
    print("procs %d: copy GB/s per proc %.1f total %.1f" % (n, np.mean(r), np.sum(r)), flush=True)
P
} > gpurun_out/r02/host_topology.txt 2>&1
tail -40 gpurun_out/r02/host_topology.txt
