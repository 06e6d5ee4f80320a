#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check_2.json 2> gpurun_out/dist_check_2.err
echo "exit $?" >> gpurun_out/dist_check_2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 1 --warmup 1 --dofs 250000 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "exit $?" >> gpurun_out/bench_2gpu.err
timeout 600 python bench.py --gpus 1 --steps 1 --warmup 1 --dofs 250000 --spmv-dofs 0 --no-cpu-baseline > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err
cat gpurun_out/dist_check_2.json | cut -c1-2000; tail -5 gpurun_out/dist_check_2.err | cut -c1-600
cat gpurun_out/bench_2gpu.json | cut -c1-1500; tail -3 gpurun_out/bench_2gpu.err | cut -c1-600
cat gpurun_out/bench_1gpu.json | cut -c1-1500
