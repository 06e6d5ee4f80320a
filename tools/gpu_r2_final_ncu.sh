#!/bin/bash
# round 2 (ONE GPU): ncu --set full of the hot kernels on the DEFAULT workload's operator (8M DoF) for roofline.traffic
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python tools/ncu_kernels.py --dofs 8000000 --reps 1"
timeout 300 $CMD > gpurun_out/r2_ncu_kernels_8M_plain.json 2> gpurun_out/r2_ncu_kernels_8M_plain.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'sell_kernel|multi_dot_kernel|multi_axpy_kernel|basis_rotate_dmma' -c 8 -o gpurun_out/r2_prof_hot_8M $CMD > gpurun_out/r2_ncu_hot_8M.log 2>&1
tail -3 gpurun_out/r2_ncu_hot_8M.log; cat gpurun_out/r2_ncu_kernels_8M_plain.json
ncu -i gpurun_out/r2_prof_hot_8M.ncu-rep --page raw --csv > gpurun_out/r2_prof_hot_8M_raw.csv 2>/dev/null; wc -c gpurun_out/r2_prof_hot_8M_raw.csv
