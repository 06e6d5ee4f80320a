#!/bin/bash
N=${1:-2}
DOFS=${2:-250000}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check_$N.json 2> gpurun_out/dist_check_$N.err
echo "exit $?" >> gpurun_out/dist_check_$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 1 --warmup 1 --dofs $DOFS --no-cpu-baseline > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "exit $?" >> gpurun_out/bench_${N}gpu.err
cat gpurun_out/dist_check_$N.json | cut -c1-1200; grep -v "Warn\|sparse_coo" gpurun_out/dist_check_$N.err | tail -4 | cut -c1-300
cat gpurun_out/bench_${N}gpu.json | cut -c1-1800; grep -v "Warn\|sparse_coo" gpurun_out/bench_${N}gpu.err | tail -4 | cut -c1-300
