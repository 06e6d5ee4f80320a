#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# HX_TEST_EXPERIMENTAL=1 also runs the tests of the switches that are off by default
timeout 900 env HX_TEST_EXPERIMENTAL=1 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout=600 --maxfail=10 --durations=8 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|Error" gpurun_out/pytest_gpu.log | head -12
timeout 500 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "exit $?" >> gpurun_out/bench_default.err
python - <<PY
import json
b=json.loads(open('gpurun_out/bench_default.json').read().strip().split('\n')[-1])
for k in ('value','e2e','gpu_launches','solver_stats','omega','roofline','iteration','cpu_baseline'): print(k, b.get(k))
PY
