#!/bin/bash
# round 2 (ONE GPU): 8M-DoF step (candidate default workload), then the ncu evidence
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py --dofs 8000000 --steps 2 --warmup 1 --record-dofs 0 --anchor-dofs 0 > gpurun_out/r2_bench_8M_n1.json 2> gpurun_out/r2_bench_8M_n1.err
python - <<PY
import json
b=json.loads(open('gpurun_out/r2_bench_8M_n1.json').read().strip().split('\n')[-1])
for k in ('value','solver_stats','omega','phases','config'): print('8M', k, b.get(k))
PY
bash tools/gpu_r2_ncu.sh 5000000
