#!/bin/bash
# round 2 (ONE GPU): whole suite on the new coarse levels / SELL transfers, 5M and 1M bench lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout=800 --durations=6 > gpurun_out/r2_pytest_6.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_pytest_6.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2_pytest_6.log | head -8
for D in 5000000 1000000; do
timeout 600 python bench.py --dofs $D --steps 2 --warmup 1 --record-dofs 0 --anchor-dofs 0 > gpurun_out/r2_bench_${D}_c.json 2> gpurun_out/r2_bench_${D}_c.err
python - <<PY
import json
b=json.loads(open('gpurun_out/r2_bench_${D}_c.json').read().strip().split('\n')[-1])
for k in ('value','solver_stats','omega','omega_check','phases','iteration'): print($D, k, b.get(k))
PY
done
HX_AMG_SELL_TRANSFER=0 timeout 600 python bench.py --dofs 5000000 --steps 1 --warmup 1 --record-dofs 0 --anchor-dofs 0 --no-phases > gpurun_out/r2_bench_5M_csrtransfer.json 2> gpurun_out/r2_bench_5M_csrtransfer.err
python - <<PY
import json
b=json.loads(open('gpurun_out/r2_bench_5M_csrtransfer.json').read().strip().split('\n')[-1])
for k in ('value','solver_stats','iteration'): print('csr-transfer', k, b.get(k))
PY
