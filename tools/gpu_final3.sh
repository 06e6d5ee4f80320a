#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 --maxfail=10 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "exit $?" >> gpurun_out/bench_default.err
cat gpurun_out/bench_default.json | cut -c1-200; python - <<PY
import json
b=json.loads(open('gpurun_out/bench_default.json').read().strip().split('\n')[-1])
for k in ('value','e2e','gpu_launches','solver_stats','omega','roofline','spmv_10m','clocks'): print(k, b[k])
print('cpu', b['cpu_baseline']['value'], b['cpu_baseline']['sample_seconds'])
PY
timeout 600 python bench.py --dofs 10000000 --steps 1 --warmup 0 --spmv-dofs 0 --no-cpu-baseline > gpurun_out/bench_10M.json 2> gpurun_out/bench_10M.err
echo "exit $?" >> gpurun_out/bench_10M.err
python - <<PY
import json
b=json.loads(open('gpurun_out/bench_10M.json').read().strip().split('\n')[-1])
for k in ('value','e2e','gpu_launches','solver_stats','omega','clocks'): print('10M', k, b[k])
PY
