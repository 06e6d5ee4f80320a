#!/bin/bash
# ncu evidence for profiles/ (one GPU; run only after the plain commands below exited 0):
#  1. `--set full` capture of the SELL SpMV kernel on the 1 M-DoF operator (2 launches after warm-up)
#  2. launch list (gpu__time_duration.sum) of one preconditioned inner solve, bracketed by
#     cudaProfilerStart/Stop inside tools/profile_solve.py (never skip tens of thousands of launches
#     with -s: ncu replays every skipped kernel's launch overhead and the box time is charged)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python tools/spmv_bench.py --dofs 1000000 --launches 20 --warmup 5"
timeout 300 $CMD > gpurun_out/spmv_bench_1m.json 2> gpurun_out/spmv_bench.err && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sell_kernel' -s 30 -c 2 -o gpurun_out/prof_sell_1m $CMD > gpurun_out/ncu_sell_1m.log 2>&1
tail -2 gpurun_out/ncu_sell_1m.log
CMD="python tools/profile_solve.py --dofs 1000000 --agg 16"
timeout 600 $CMD > gpurun_out/profile_solve_1M.json 2> gpurun_out/profile_solve.err && \
HX_AMG_GRAPH=0 timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_solve_1M.csv $CMD > gpurun_out/ncu_solve.log 2>&1
cat gpurun_out/profile_solve_1M.json; wc -l gpurun_out/launches_solve_1M.csv; tail -2 gpurun_out/ncu_solve.log
