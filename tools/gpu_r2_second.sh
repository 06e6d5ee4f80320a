#!/bin/bash
# round 2, second GPU job: suite on the new defaults + device colouring, bench with phases / anchor / 10M record
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout=600 --durations=8 > gpurun_out/r2_pytest_2.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_pytest_2.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2_pytest_2.log | head -12
timeout 900 python bench.py --dofs 1000000 --steps 2 --warmup 1 > gpurun_out/r2_bench_1M.json 2> gpurun_out/r2_bench_1M.err
echo "exit $?" >> gpurun_out/r2_bench_1M.err; tail -3 gpurun_out/r2_bench_1M.err
timeout 600 python bench.py --dofs 10000000 --steps 1 --warmup 0 --record-dofs 0 --anchor-dofs 0 > gpurun_out/r2_bench_10M_phases.json 2> gpurun_out/r2_bench_10M_phases.err
echo "exit $?" >> gpurun_out/r2_bench_10M_phases.err; tail -3 gpurun_out/r2_bench_10M_phases.err
python - <<PY
import json
for f in ("r2_bench_1M","r2_bench_10M_phases"):
    try:
        b=json.loads(open('gpurun_out/%s.json'%f).read().strip().split('\n')[-1])
        for k in ('value','e2e','device_span_s','gpu_launches','solver_stats','omega','phases','roofline','iteration','record_10m','anchor','clocks'): print(f, k, b.get(k))
    except Exception as e: print(f, 'failed', e)
PY
