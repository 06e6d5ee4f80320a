#!/bin/bash
# round 2 (ONE GPU): fused cycle tail -- suite, then 1M / 5M steps with and without it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout=800 --durations=6 > gpurun_out/r2_pytest_8.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_pytest_8.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2_pytest_8.log | head -8
run() {  # dofs, name, env...
  D=$1; name=$2; shift 2
  env "$@" timeout 600 python bench.py --dofs $D --steps 1 --warmup 1 --record-dofs 0 --anchor-dofs 0 --no-phases \
      > gpurun_out/tail_${name}_$D.json 2> gpurun_out/tail_${name}_$D.err
  python - <<PY
import json
try:
    b = json.loads(open('gpurun_out/tail_${name}_$D.json').read().strip().split('\n')[-1])
    it = b.get('iteration') or {}
    print('${name}', $D, 'value', b['value'], 'its', b['solver_stats']['inner_iterations'], 't_inner', b['solver_stats']['t_inner'],
          'cycle', it.get('multigrid_cycle_us'), it.get('cycle_visit_us'), b['omega'])
except Exception as e:
    print('${name}', 'failed', e)
PY
  grep -v "Warn\|sparse_coo" gpurun_out/tail_${name}_$D.err | tail -2 | cut -c1-300
}
run 1000000 fused HX_AMG_TAIL=1
run 1000000 unfused HX_AMG_TAIL=0
run 250000 fused HX_AMG_TAIL=1
run 250000 unfused HX_AMG_TAIL=0
run 5000000 fused HX_AMG_TAIL=1
