#!/bin/bash
# round 2, fourth GPU job (ONE GPU): reference drivers unchanged, cycle shapes at 10M, 5M step for the default workload
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_drivers.py -x -q -p no:cacheprovider --timeout=600 --durations=4 > gpurun_out/r2_pytest_drivers.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_pytest_drivers.log
tail -30 gpurun_out/r2_pytest_drivers.log | cut -c1-600
run() {  # dofs, name, env...
  D=$1; name=$2; shift 2
  env "$@" timeout 600 python bench.py --dofs $D --steps 1 --warmup 0 --record-dofs 0 --anchor-dofs 0 --no-phases \
      > gpurun_out/shape_${name}_$D.json 2> gpurun_out/shape_${name}_$D.err
  python - <<PY
import json
try:
    b = json.loads(open('gpurun_out/shape_${name}_$D.json').read().strip().split('\n')[-1])
    print('${name}', $D, 'value', b['value'], b['solver_stats'], (b.get('iteration') or {}).get('multigrid_cycle_us'), b['omega'])
except Exception as e:
    print('${name}', 'failed', e)
PY
}
run 10000000 w1to1 HX_AMG_WCYCLE=1:1
run 10000000 w1to2 HX_AMG_WCYCLE=1:2
run 10000000 w2to2 HX_AMG_WCYCLE=2:2
timeout 600 python bench.py --dofs 5000000 --steps 2 --warmup 1 --record-dofs 0 --anchor-dofs 0 > gpurun_out/r2_bench_5M_n1.json 2> gpurun_out/r2_bench_5M_n1.err
python - <<PY
import json
b=json.loads(open('gpurun_out/r2_bench_5M_n1.json').read().strip().split('\n')[-1])
for k in ('value','e2e','solver_stats','omega','phases','config'): print('5M', k, b.get(k))
PY
