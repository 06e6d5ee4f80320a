#!/bin/bash
# round 2, first GPU job: whole GPU suite on the fixed inner solve, then A/B of the CPU-validated switches
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 --durations=8 > gpurun_out/r2_pytest_1.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_pytest_1.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2_pytest_1.log | head -12
run() {  # dofs, warmup, name, env...
  D=$1; W=$2; name=$3; shift 3
  env "$@" timeout 600 python bench.py --dofs $D --steps 1 --warmup $W --spmv-dofs 0 --no-cpu-baseline \
      > gpurun_out/ab_${name}_$D.json 2> gpurun_out/ab_${name}_$D.err
  python - <<PY
import json
try:
    b = json.loads(open('gpurun_out/ab_${name}_$D.json').read().strip().split('\n')[-1])
    print('${name}', $D, 'value', b['value'], 'e2e', b['e2e']['value'], b['solver_stats'], b.get('iteration'), b['omega'])
except Exception as e:
    print('${name}', 'failed', e)
PY
}
run 1000000 1 cgs1 HX_GMRES_ORTH=cgs1
run 1000000 1 cgs2 HX_GMRES_ORTH=cgs2
run 1000000 1 cheb HX_AMG_SMOOTHER=chebyshev
run 1000000 1 relax HX_INNER_RELAX=1
run 1000000 1 cheb_relax HX_AMG_SMOOTHER=chebyshev HX_INNER_RELAX=1
run 1000000 1 cheb_relax_w2 HX_AMG_SMOOTHER=chebyshev HX_INNER_RELAX=1 HX_AMG_WCYCLE=2
run 1000000 1 cheb_relax_w1 HX_AMG_SMOOTHER=chebyshev HX_INNER_RELAX=1 HX_AMG_WCYCLE=1
run 10000000 0 cgs1 HX_GMRES_ORTH=cgs1
run 10000000 0 cheb_relax HX_AMG_SMOOTHER=chebyshev HX_INNER_RELAX=1
run 10000000 0 cheb_relax_w1 HX_AMG_SMOOTHER=chebyshev HX_INNER_RELAX=1 HX_AMG_WCYCLE=1
