#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check_2.json 2> gpurun_out/dist_check_2.err
echo "exit $?" >> gpurun_out/dist_check_2.err
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 1 --warmup 0 --no-cpu-baseline --spmv-dofs 0 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "exit $?" >> gpurun_out/bench_2gpu.err
cat gpurun_out/dist_check_2.json | cut -c1-1000; grep -v "Warn\|sparse_coo" gpurun_out/dist_check_2.err | tail -4 | cut -c1-300
cat gpurun_out/bench_2gpu.json | cut -c1-1800; grep -v "Warn\|sparse_coo" gpurun_out/bench_2gpu.err | tail -4 | cut -c1-300
