#!/bin/bash
# round 2 ncu evidence (ONE GPU).  Each ncu run follows the same command exiting 0 without ncu.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
D=${1:-5000000}
CMD="python tools/ncu_kernels.py --dofs $D --reps 2"
timeout 300 $CMD > gpurun_out/r2_ncu_kernels_plain.json 2> gpurun_out/r2_ncu_kernels_plain.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'sell_kernel|multi_dot_kernel|multi_axpy_kernel|basis_rotate_dmma' -c 24 -o gpurun_out/r2_prof_hot $CMD > gpurun_out/r2_ncu_hot.log 2>&1
tail -3 gpurun_out/r2_ncu_hot.log; cat gpurun_out/r2_ncu_kernels_plain.json
CMD2="python tools/profile_solve.py --dofs 1000000 --agg 16"
HX_AMG_GRAPH=0 timeout 300 $CMD2 > gpurun_out/r2_profile_solve_1M.json 2> gpurun_out/r2_profile_solve.err && \
HX_AMG_GRAPH=0 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_inner_solve_1M.csv $CMD2 > gpurun_out/r2_ncu_solve.log 2>&1
cat gpurun_out/r2_profile_solve_1M.json | cut -c1-600; wc -l gpurun_out/r2_launches_inner_solve_1M.csv; tail -2 gpurun_out/r2_ncu_solve.log
ls -la gpurun_out/r2_prof_hot.ncu-rep
