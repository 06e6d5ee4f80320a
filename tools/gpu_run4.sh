#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python tools/profile_solve.py --dofs 1000000"
timeout 600 $CMD > gpurun_out/profile_solve_1M.json 2> gpurun_out/profile_solve.err && \
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_solve_1M.csv $CMD > gpurun_out/ncu_solve.log 2>&1
cat gpurun_out/profile_solve_1M.json; wc -l gpurun_out/launches_solve_1M.csv; tail -2 gpurun_out/ncu_solve.log
