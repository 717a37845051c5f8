#!/bin/bash
# N=8 and N=4 bench lines (device value + e2e through the plugin)
mkdir -p gpurun_out/r02
for n in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
    bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline --no-visit-line > gpurun_out/r02/bench_n$n.json 2> gpurun_out/r02/bench_n$n.err
  echo "N=$n rc=$?"; tail -c 300 gpurun_out/r02/bench_n$n.err
done
