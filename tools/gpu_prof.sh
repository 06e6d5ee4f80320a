#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python tools/spmv_bench.py --dofs 10000000 --launches 20 --warmup 5"
$CMD > gpurun_out/spmv_bench_10m.json 2> gpurun_out/spmv_bench.err && \
ncu --set full --clock-control none --import-source on -k regex:'spmv_(csr|sell)_kernel' -s 40 -c 2 -o gpurun_out/prof_spmv_csr8 $CMD > gpurun_out/ncu_spmv.log 2>&1
# SELL kernel capture: it is launched after the 4 CSR variants (4*(5+20)=100 launches)
ncu --set full --clock-control none --import-source on -k regex:'spmv_sell_kernel' -s 10 -c 2 -o gpurun_out/prof_spmv_sell $CMD > gpurun_out/ncu_sell.log 2>&1
cat gpurun_out/spmv_bench_10m.json; tail -3 gpurun_out/ncu_spmv.log; tail -3 gpurun_out/ncu_sell.log
BCMD="python bench.py --steps 1 --warmup 0 --dofs 100000 --spmv-dofs 0 --no-cpu-baseline"
$BCMD > gpurun_out/bench_for_ncu.json 2> gpurun_out/bench_for_ncu.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 60000 -c 6000 --csv --log-file gpurun_out/launches_step.csv $BCMD > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/ncu_step.log; wc -l gpurun_out/launches_step.csv
