#!/bin/bash
# round-2 GPU job A: gpu tests, bench with / without the compiled XyToV, occupancy variants, ncu capture of k_pool_step
mkdir -p gpurun_out/r02
python -m pytest tests -m gpu -x -q > gpurun_out/r02/pytest_gpu_2.log 2>&1; tail -5 gpurun_out/r02/pytest_gpu_2.log
python bench.py --no-visit-line --no-cpu-baseline > gpurun_out/r02/bench_xyv.json 2> gpurun_out/r02/bench_xyv.err
B2_XYTOV_EXACT=1 python bench.py --no-visit-line --no-cpu-baseline > gpurun_out/r02/bench_xyv_exact.json 2>> gpurun_out/r02/bench_xyv.err
for occ in 3 4; do B2_POOL_OCC=$occ python bench.py --no-visit-line --no-cpu-baseline --steps 5 > gpurun_out/r02/bench_xyv_occ$occ.json 2>> gpurun_out/r02/bench_xyv.err; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pool_step -s 2 -c 1 -o gpurun_out/r02/prof_pool_xyv -f python bench.py --no-visit-line --no-cpu-baseline --steps 2 --warmup 3 > gpurun_out/r02/ncu_pool.log 2>&1; tail -2 gpurun_out/r02/ncu_pool.log
python - <<PY
import json
for f in ["bench_xyv","bench_xyv_exact","bench_xyv_occ3","bench_xyv_occ4"]:
    try:
        d=json.load(open("gpurun_out/r02/%s.json"%f)); print(f, d["value"], d["ms_per_step"], d["roofline_fp64"]["frac"], d["breakdown_ms"], d["e2e"]["value"])
    except Exception as e: print(f, "ERR", e)
PY
