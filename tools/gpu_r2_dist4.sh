#!/bin/bash
# round 2: N GPUs, bench workload, distributed-level threshold A/B
N=${1:-4}
DOFS=${2:-5000000}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 2 --warmup 1 --dofs $DOFS > gpurun_out/r2_bench_${N}gpu_$name.json 2> gpurun_out/r2_bench_${N}gpu_$name.err
  echo "exit $?" >> gpurun_out/r2_bench_${N}gpu_$name.err
  python - <<PY
import json
try:
    b=json.loads(open('gpurun_out/r2_bench_${N}gpu_$name.json').read().strip().split('\n')[-1])
    for k in ('value','e2e','solver_stats','omega_check','phases'): print('$name', k, b[k])
except Exception as e: print('$name bench parse failed', e)
PY
  grep -v "Warn\|sparse_coo\|^\*\*\*\|OMP_NUM" gpurun_out/r2_bench_${N}gpu_$name.err | tail -5 | cut -c1-400
}
run min2000 HX_DIST_MIN_ROWS=2000
run min20000 HX_DIST_MIN_ROWS=20000
