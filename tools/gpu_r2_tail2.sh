#!/bin/bash
# round 2 (ONE GPU): fused cycle tail with 4 CTAs per SM -- with / without at 1M and 8M DoF
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -p no:cacheprovider -k "fused_tail" > gpurun_out/r2_pytest_tail2.log 2>&1; tail -2 gpurun_out/r2_pytest_tail2.log
run() {  # dofs, name, env...
  D=$1; name=$2; shift 2
  env "$@" timeout 600 python bench.py --dofs $D --steps 1 --warmup 1 --record-dofs 0 --anchor-dofs 0 --no-phases \
      > gpurun_out/tail2_${name}_$D.json 2> gpurun_out/tail2_${name}_$D.err
  python - <<PY
import json
try:
    b = json.loads(open('gpurun_out/tail2_${name}_$D.json').read().strip().split('\n')[-1])
    it = b.get('iteration') or {}
    print('${name}', $D, 'value', b['value'], 'its', b['solver_stats']['inner_iterations'], 't_inner', b['solver_stats']['t_inner'],
          'cycle', it.get('multigrid_cycle_us'), it.get('cycle_visit_us'))
except Exception as e:
    print('${name}', 'failed', e)
PY
}
run 1000000 fused HX_AMG_TAIL=1
run 1000000 unfused HX_AMG_TAIL=0
run 8000000 fused HX_AMG_TAIL=1
