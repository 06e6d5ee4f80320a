#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for d in 1000000; do
timeout 1200 python bench.py --steps 1 --warmup 0 --dofs $d --spmv-dofs 10000000 --no-cpu-baseline > gpurun_out/bench_$d.json 2> gpurun_out/bench_$d.err
echo "exit $?" >> gpurun_out/bench_$d.err
done
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider -k "edge_cases" > gpurun_out/pytest_edge.log 2>&1
tail -3 gpurun_out/pytest_edge.log
cat gpurun_out/bench_1000000.json | cut -c1-2500; tail -3 gpurun_out/bench_1000000.err
