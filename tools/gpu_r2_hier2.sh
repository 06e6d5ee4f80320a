#!/bin/bash
# round 2 (ONE GPU): around the best coarse-level setting of gpu_r2_hier.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
D=${1:-5000000}
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --dofs $D --steps 1 --warmup 0 --record-dofs 0 --anchor-dofs 0 --no-phases \
      > gpurun_out/hier_${name}_$D.json 2> gpurun_out/hier_${name}_$D.err
  python - <<PY
import json
try:
    b = json.loads(open('gpurun_out/hier_${name}_$D.json').read().strip().split('\n')[-1])
    it = b.get('iteration') or {}
    print('${name}', $D, 'value', b['value'], 'its', b['solver_stats']['inner_iterations'], 'solves', b['solver_stats']['inner_solves'], 't_inner', b['solver_stats']['t_inner'], 'setup', b['solver_stats']['t_amg_setup'],
          'cycle', it.get('multigrid_cycle_us'), it.get('cycle_visit_us'), it.get('amg_levels'), it.get('amg_level_nnz'), b['omega'])
except Exception as e:
    print('${name}', 'failed', e)
PY
}
run warm HX_AMG_AGG_COARSE=32 HX_AMG_PFILTER=0.1
run a32_pf01_w1 HX_AMG_AGG_COARSE=32 HX_AMG_PFILTER=0.1 HX_AMG_WCYCLE=1
run a32_pf01_w12 HX_AMG_AGG_COARSE=32 HX_AMG_PFILTER=0.1 HX_AMG_WCYCLE=1:2
run a32_pf015 HX_AMG_AGG_COARSE=32 HX_AMG_PFILTER=0.15
run a32_pf02 HX_AMG_AGG_COARSE=32 HX_AMG_PFILTER=0.2
run a24_pf01 HX_AMG_AGG_COARSE=24 HX_AMG_PFILTER=0.1
run a48_pf01 HX_AMG_AGG_COARSE=48 HX_AMG_PFILTER=0.1
