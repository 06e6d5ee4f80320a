#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/amg_sweep.jsonl
for cfg in "2 8" "1 8" "2 16" "1 16" "2 27" "1 27" "3 8"; do
  set -- $cfg
  timeout 300 python tools/profile_solve.py --dofs 1000000 --nu $1 --agg $2 2>/dev/null | tail -1 >> gpurun_out/amg_sweep.jsonl
done
cat gpurun_out/amg_sweep.jsonl | cut -c1-400
