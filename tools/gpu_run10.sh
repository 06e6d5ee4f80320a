#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 --maxfail=10 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
: > gpurun_out/amg_sweep2.jsonl
for cfg in "2 16 single" "2 16 double" "2 8 single" "2 8 double" "1 16 single"; do
  set -- $cfg
  timeout 300 python tools/profile_solve.py --dofs 1000000 --nu $1 --agg $2 --precision $3 2>/dev/null | tail -1 >> gpurun_out/amg_sweep2.jsonl
done
timeout 300 python tools/profile_solve.py --dofs 250000 --precision single 2>/dev/null | tail -1 >> gpurun_out/amg_sweep2.jsonl
cat gpurun_out/amg_sweep2.jsonl | cut -c1-330
timeout 900 python bench.py --steps 2 --warmup 1 --dofs 1000000 --spmv-dofs 0 --no-cpu-baseline > gpurun_out/bench_1M.json 2> gpurun_out/bench_1M.err
echo "exit $?" >> gpurun_out/bench_1M.err
cat gpurun_out/bench_1M.json | cut -c1-1300
