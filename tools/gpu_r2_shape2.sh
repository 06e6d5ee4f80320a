#!/bin/bash
# round 2 (ONE GPU): cycle shape on the new coarse levels (aggregates of 32 below level 0, prolongator filter 0.1)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # dofs, name, env...
  D=$1; name=$2; shift 2
  env "$@" timeout 600 python bench.py --dofs $D --steps 1 --warmup 1 --record-dofs 0 --anchor-dofs 0 --no-phases \
      > gpurun_out/shape2_${name}_$D.json 2> gpurun_out/shape2_${name}_$D.err
  python - <<PY
import json
try:
    b = json.loads(open('gpurun_out/shape2_${name}_$D.json').read().strip().split('\n')[-1])
    it = b.get('iteration') or {}
    print('${name}', $D, 'value', b['value'], 'its', b['solver_stats']['inner_iterations'], 'solves', b['solver_stats']['inner_solves'], 't_inner', b['solver_stats']['t_inner'], 'setup', b['solver_stats']['t_amg_setup'],
          'cycle', it.get('multigrid_cycle_us'), it.get('cycle_visit_us'), it.get('amg_levels'), b['omega'])
except Exception as e:
    print('${name}', 'failed', e)
PY
}
run 1000000 v HX_AMG_WCYCLE=off
run 1000000 w1 HX_AMG_WCYCLE=1
run 1000000 w22 HX_AMG_WCYCLE=2:2
run 250000 v HX_AMG_WCYCLE=off
run 250000 w1 HX_AMG_WCYCLE=1
run 10000000 w1 HX_AMG_WCYCLE=1
run 10000000 w22 HX_AMG_WCYCLE=2:2
