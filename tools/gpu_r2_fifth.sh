#!/bin/bash
# round 2, fifth GPU job (ONE GPU): whole suite, DMMA vs FMA rotation, 5M phases with the multigrid set-up split
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout=800 --durations=6 > gpurun_out/r2_pytest_5.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_pytest_5.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2_pytest_5.log | head -8
timeout 300 python tools/rotate_bench.py > gpurun_out/r2_rotate_bench.json 2> gpurun_out/r2_rotate_bench.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_rotate_bench.json').read().strip().split('\n')[-1])
    for c in d['cases']: print(c['n'], c['m'], c['k'], 'fma', c['fma'], 'dmma', c['dmma'], 'speedup', c['dmma_speedup'])
except Exception as e: print('rotate failed', e)
PY
tail -3 gpurun_out/r2_rotate_bench.err | cut -c1-300
timeout 600 python bench.py --dofs 5000000 --steps 1 --warmup 1 --record-dofs 0 --anchor-dofs 0 > gpurun_out/r2_bench_5M_n1_b.json 2> gpurun_out/r2_bench_5M_n1_b.err
python - <<PY
import json
b=json.loads(open('gpurun_out/r2_bench_5M_n1_b.json').read().strip().split('\n')[-1])
for k in ('value','solver_stats','omega','phases','iteration'): print('5M', k, b.get(k))
PY
