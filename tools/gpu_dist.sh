#!/bin/bash
N=${1:-8}
DOFS=${2:-1000000}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check_$N.json 2> gpurun_out/dist_check_$N.err
echo "exit $?" >> gpurun_out/dist_check_$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 1 --warmup 1 --dofs $DOFS --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "exit $?" >> gpurun_out/bench_${N}gpu.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/dist_check_$N.json').read().strip().split('\n')[-1])
    for k in ('rijke3d','annulus'): print(k, {q:d[0][k][q] for q in ('seconds','omega','max_abs_diff_vs_log')}, d[0][k]['stats'])
    print('annulus rel', d[0]['annulus'].get('rel_diff_vs_eigenvalues_dir'))
except Exception as e: print('dist_check parse failed', e)
try:
    b=json.loads(open('gpurun_out/bench_${N}gpu.json').read().strip().split('\n')[-1])
    for k in ('value','e2e','gpu_launches','solver_stats','omega'): print(k, b[k])
except Exception as e: print('bench parse failed', e)
PY
grep -v "Warn\|sparse_coo" gpurun_out/dist_check_$N.err | tail -3 | cut -c1-300; grep -v "Warn\|sparse_coo" gpurun_out/bench_${N}gpu.err | tail -3 | cut -c1-300
