#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 --maxfail=10 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --dofs 10000000 --steps 1 --warmup 0 --spmv-dofs 0 --no-cpu-baseline > gpurun_out/bench_10M.json 2> gpurun_out/bench_10M.err
echo "exit $?" >> gpurun_out/bench_10M.err
cat gpurun_out/bench_10M.json | cut -c1-300; tail -c 1200 gpurun_out/bench_10M.json; grep -v "Warn\|sparse_coo" gpurun_out/bench_10M.err | tail -3 | cut -c1-300
