#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 python bench.py --dofs 10000000 --steps 1 --warmup 0 --spmv-dofs 0 --no-cpu-baseline > gpurun_out/bench_10M.json 2> gpurun_out/bench_10M.err
echo "exit $?" >> gpurun_out/bench_10M.err
python - <<PY
import json
b=json.loads(open('gpurun_out/bench_10M.json').read().strip().split('\n')[-1])
for k in ('value','e2e','gpu_launches','solver_stats','omega','clocks'): print('10M', k, b[k])
PY
