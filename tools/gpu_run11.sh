#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 --maxfail=10 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/profile_solve.py --dofs 250000 --precision single 2>/dev/null | tail -1 | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 2 --warmup 1 --dofs 1000000 --spmv-dofs 0 > gpurun_out/bench_1M.json 2> gpurun_out/bench_1M.err
echo "exit $?" >> gpurun_out/bench_1M.err
cat gpurun_out/bench_1M.json | cut -c1-1300; tail -c 900 gpurun_out/bench_1M.json
