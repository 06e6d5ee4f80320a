#!/bin/bash
# round 2: N real GPUs -- goldens through the peer transport, then the bench workload with phases
N=${1:-2}
DOFS=${2:-5000000}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/r2_dist_check_$N.json 2> gpurun_out/r2_dist_check_$N.err
echo "exit $?" >> gpurun_out/r2_dist_check_$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 2 --warmup 1 --dofs $DOFS > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
echo "exit $?" >> gpurun_out/r2_bench_${N}gpu.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_dist_check_$N.json').read().strip().split('\n')[-1])
    print('unit', d[0].get('unit'))
    for k in ('rijke3d','annulus'): print(k, {q:d[0][k][q] for q in ('seconds','omega','max_abs_diff_vs_log','distributed_levels','cycle_in_graph')}, d[0][k]['stats'])
    print('annulus rel', d[0]['annulus'].get('rel_diff_vs_eigenvalues_dir'))
except Exception as e: print('dist_check parse failed', e)
try:
    b=json.loads(open('gpurun_out/r2_bench_${N}gpu.json').read().strip().split('\n')[-1])
    for k in ('value','e2e','gpu_launches','solver_stats','omega','omega_check','phases','roofline'): print(k, b[k])
except Exception as e: print('bench parse failed', e)
PY
grep -v "Warn\|sparse_coo\|^\*\*\*\|OMP_NUM" gpurun_out/r2_dist_check_$N.err | tail -8 | cut -c1-400; grep -v "Warn\|sparse_coo\|^\*\*\*\|OMP_NUM" gpurun_out/r2_bench_${N}gpu.err | tail -8 | cut -c1-400
