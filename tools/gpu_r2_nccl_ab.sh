#!/bin/bash
# round 2 (2 GPUs): the same step with the NCCL data path (send/recv halo + all_reduce from Python, cycle not in a graph)
# against the peer-memory kernels (default) -- what the own kernels buy over the library baseline
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-2}
D=${2:-8000000}
HX_DIST_TRANSPORT=nccl timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 2 --warmup 1 --dofs $D --no-phases > gpurun_out/r2_nccl_${N}gpu.json 2> gpurun_out/r2_nccl_${N}gpu.err
echo "exit $?" >> gpurun_out/r2_nccl_${N}gpu.err
python - <<PY
import json
try:
    b=json.loads(open('gpurun_out/r2_nccl_${N}gpu.json').read().strip().split('\n')[-1])
    for k in ('value','e2e','solver_stats','omega_check'): print('nccl', $N, k, b[k])
except Exception as e: print('nccl bench parse failed', e)
PY
grep -v "Warn\|sparse_coo\|^\*\*\*\|OMP_NUM\|pmax" gpurun_out/r2_nccl_${N}gpu.err | tail -5 | cut -c1-400
