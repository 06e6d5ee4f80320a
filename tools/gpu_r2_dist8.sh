#!/bin/bash
# round 2: 8 GPUs -- goldens (P1 + P2), then the bench workload at two sizes
N=${1:-8}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/r2_dist_check_$N.json 2> gpurun_out/r2_dist_check_$N.err
echo "exit $?" >> gpurun_out/r2_dist_check_$N.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_dist_check_$N.json').read().strip().split('\n')[-1])
    print('unit', d[0].get('unit'))
    for k in ('rijke3d','annulus'): print(k, {q:d[0][k][q] for q in ('seconds','omega','max_abs_diff_vs_log','distributed_levels','cycle_in_graph')})
    print('annulus rel', d[0]['annulus'].get('rel_diff_vs_eigenvalues_dir'))
    print('p2', d[0].get('rijke3d_p2'))
except Exception as e: print('dist_check parse failed', e)
PY
grep -v "Warn\|sparse_coo\|^\*\*\*\|OMP_NUM\|pmax" gpurun_out/r2_dist_check_$N.err | tail -6 | cut -c1-500
for DOFS in 8000000 5000000; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 2 --warmup 1 --dofs $DOFS > gpurun_out/r2_bench_${N}gpu_$DOFS.json 2> gpurun_out/r2_bench_${N}gpu_$DOFS.err
echo "exit $?" >> gpurun_out/r2_bench_${N}gpu_$DOFS.err
python - <<PY
import json
try:
    b=json.loads(open('gpurun_out/r2_bench_${N}gpu_$DOFS.json').read().strip().split('\n')[-1])
    for k in ('value','e2e','solver_stats','omega','omega_check','phases'): print($DOFS, k, b[k])
except Exception as e: print($DOFS, 'bench parse failed', e)
PY
grep -v "Warn\|sparse_coo\|^\*\*\*\|OMP_NUM\|pmax" gpurun_out/r2_bench_${N}gpu_$DOFS.err | tail -5 | cut -c1-400
done
