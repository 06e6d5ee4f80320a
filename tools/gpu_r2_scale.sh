#!/bin/bash
# round 2: the default workload on N GPUs (as the driver's scaling run does, with fewer steps)
N=${1:-8}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 1 > gpurun_out/r2_scale_${N}gpu.json 2> gpurun_out/r2_scale_${N}gpu.err
echo "exit $?" >> gpurun_out/r2_scale_${N}gpu.err
python - <<PY
import json
try:
    b=json.loads(open('gpurun_out/r2_scale_${N}gpu.json').read().strip().split('\n')[-1])
    for k in ('value','e2e','solver_stats','omega_check','phases','clocks'): print($N, k, b[k])
except Exception as e: print($N, 'bench parse failed', e)
PY
grep -v "Warn\|sparse_coo\|^\*\*\*\|OMP_NUM\|pmax" gpurun_out/r2_scale_${N}gpu.err | tail -5 | cut -c1-400
