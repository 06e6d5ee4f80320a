#!/bin/bash
# first GPU pass: parity tests (all, not -x) + a small bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 --maxfail=40 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 1 --warmup 1 --dofs 100000 --spmv-dofs 2000000 --no-cpu-baseline > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err
echo "bench exit $?" >> gpurun_out/bench_small.err
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_small.json | cut -c1-1500
