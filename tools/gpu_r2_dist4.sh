#!/bin/bash
# round 2: N GPUs -- goldens (P1 + P2) through the peer transport, then the bench workload with two cycle shapes
N=${1:-4}
DOFS=${2:-5000000}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/r2_dist_check_$N.json 2> gpurun_out/r2_dist_check_$N.err
echo "exit $?" >> gpurun_out/r2_dist_check_$N.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_dist_check_$N.json').read().strip().split('\n')[-1])
    print('unit', d[0].get('unit'))
    for k in ('rijke3d','annulus'): print(k, {q:d[0][k][q] for q in ('seconds','omega','max_abs_diff_vs_log','distributed_levels','cycle_in_graph')}, d[0][k]['stats'])
    print('annulus rel', d[0]['annulus'].get('rel_diff_vs_eigenvalues_dir'))
    print('p2', d[0].get('rijke3d_p2'))
except Exception as e: print('dist_check parse failed', e)
PY
grep -v "Warn\|sparse_coo\|^\*\*\*\|OMP_NUM" gpurun_out/r2_dist_check_$N.err | tail -8 | cut -c1-600
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
run w1 HX_X=0
run w22 HX_AMG_WCYCLE=2:2
