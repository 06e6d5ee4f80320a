#!/bin/bash
# A/B timings of the options that were added after the last GPU run of round 1 (all CPU-validated):
#   single-pass Gram-Schmidt in the inner GMRES (default on; HX_GMRES_ORTH=cgs2 = previous behaviour)
#   W-cycle (HX_AMG_WCYCLE=1, default off)
#   relaxed inner tolerance in the Krylov-Schur steps (HX_INNER_RELAX=1, default off)
#   Chebyshev-root damping of the Jacobi sweeps (HX_AMG_SMOOTHER=chebyshev, default off)
# usage: bash tools/gpu_ab.sh [dofs]     (one GPU; ~1 min per line at 1 M DoF)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
D=${1:-1000000}
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 900 python bench.py --dofs $D --steps 1 --warmup 1 --spmv-dofs 0 --no-cpu-baseline \
      > gpurun_out/ab_${name}_$D.json 2> gpurun_out/ab_${name}_$D.err
  python - <<PY
import json
try:
    b = json.loads(open('gpurun_out/ab_${name}_$D.json').read().strip().split('\n')[-1])
    print('${name}', 'value', b['value'], 'e2e', b['e2e']['value'], b['solver_stats'], b.get('iteration'))
except Exception as e:
    print('${name}', 'failed', e)
PY
}
run cgs1 HX_GMRES_ORTH=cgs1
run cgs2 HX_GMRES_ORTH=cgs2
run cgs1_w1 HX_GMRES_ORTH=cgs1 HX_AMG_WCYCLE=1
run cgs1_w2 HX_GMRES_ORTH=cgs1 HX_AMG_WCYCLE=2
run cgs1_cheb HX_GMRES_ORTH=cgs1 HX_AMG_SMOOTHER=chebyshev
run cgs1_cheb_w1 HX_GMRES_ORTH=cgs1 HX_AMG_SMOOTHER=chebyshev HX_AMG_WCYCLE=1
run cgs1_relax HX_GMRES_ORTH=cgs1 HX_INNER_RELAX=1
run cgs1_cheb_relax HX_GMRES_ORTH=cgs1 HX_AMG_SMOOTHER=chebyshev HX_INNER_RELAX=1
