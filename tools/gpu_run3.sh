#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 --maxfail=10 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python tools/spmv_bench.py --dofs 10000000 --launches 30 --warmup 5 > gpurun_out/spmv_bench_10m.json 2> gpurun_out/spmv_bench.err
timeout 900 python bench.py --steps 1 --warmup 0 --dofs 1000000 --spmv-dofs 0 --no-cpu-baseline > gpurun_out/bench_1M.json 2> gpurun_out/bench_1M.err
echo "exit $?" >> gpurun_out/bench_1M.err
tail -4 gpurun_out/pytest_gpu.log; cat gpurun_out/spmv_bench_10m.json; cat gpurun_out/bench_1M.json | cut -c1-1800
