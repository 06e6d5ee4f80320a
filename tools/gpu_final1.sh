#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 --maxfail=10 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python tools/profile_solve.py --dofs 1000000 2>/dev/null | tail -1 | cut -c1-420
CMD="python tools/spmv_bench.py --dofs 1000000 --launches 20 --warmup 5"
timeout 300 $CMD > gpurun_out/spmv_bench_1m.json 2> gpurun_out/spmv_bench.err && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sell_kernel' -s 30 -c 2 -o gpurun_out/prof_sell_1m $CMD > gpurun_out/ncu_sell_1m.log 2>&1
tail -2 gpurun_out/ncu_sell_1m.log
timeout 1200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "exit $?" >> gpurun_out/bench_default.err
cat gpurun_out/bench_default.json | cut -c1-600; tail -c 1500 gpurun_out/bench_default.json
