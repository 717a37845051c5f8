#!/bin/bash
mkdir -p gpurun_out/r02
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --no-visit-line > gpurun_out/r02/bench_2gpu.json 2> gpurun_out/r02/bench_2gpu.err; tail -2 gpurun_out/r02/bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02/bench_2gpu.json").read().strip().splitlines()[-1])
print("N=2 value", d["value"], "e2e", d["e2e"]["value"], d["e2e"].get("wall_s"), "pinned", d["e2e"].get("pinned_route",{}).get("value"))
PY
nproc; free -g | head -2
