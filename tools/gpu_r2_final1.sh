#!/bin/bash
# round 2 (ONE GPU): the suite and the default bench line exactly as the driver runs it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider --timeout=800 --durations=6 > gpurun_out/r2_pytest_final.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_pytest_final.log
grep -E "passed|failed|FAILED|Error" gpurun_out/r2_pytest_final.log | head -8
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log
SECONDS=0
timeout 1700 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_default_n1.json 2> gpurun_out/r2_bench_default_n1.err
echo "bench exit $? after $SECONDS s"
python - <<PY
import json
b=json.loads(open('gpurun_out/r2_bench_default_n1.json').read().strip().split('\n')[-1])
for k in ('value','e2e','device_span_s','loop_wall_s','gpu_launches','solver_stats','omega','omega_check','phases','roofline','record_10m','anchor','clocks'): print(k, b.get(k))
PY
SECONDS=0
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err
echo "reference arm exit $? after $SECONDS s"; cut -c1-700 gpurun_out/r2_bench_reference_arm.json
