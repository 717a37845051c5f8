#!/bin/bash
mkdir -p gpurun_out/r02
python -m pytest tests/test_gpu_sensor.py -x -q 2>&1 | grep -B14 "AttributeError" | head -40
python bench.py --no-visit-line --no-cpu-baseline --steps 5 > gpurun_out/r02/bench_configs.json 2> gpurun_out/r02/bench_configs.err; tail -3 gpurun_out/r02/bench_configs.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02/bench_configs.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"].get("wall_s"))
for k,v in (d.get("configs") or {}).items():
    print(k, {a:b for a,b in v.items() if a!="what"} if isinstance(v, dict) else v)
PY
