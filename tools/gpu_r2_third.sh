#!/bin/bash
# round 2, third GPU job (ONE GPU): two ranks on cuda:0 over the peer-memory transport, per-level cycle timings at 10M
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q -p no:cacheprovider --timeout=800 --durations=4 > gpurun_out/r2_pytest_dist.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_pytest_dist.log
tail -25 gpurun_out/r2_pytest_dist.log | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/dist_check.py --same-device --no-unit --cases annulus > gpurun_out/r2_dist_same_annulus.json 2> gpurun_out/r2_dist_same_annulus.err
echo "exit $?" >> gpurun_out/r2_dist_same_annulus.err
tail -c 1500 gpurun_out/r2_dist_same_annulus.json; grep -v "Warn\|sparse_coo" gpurun_out/r2_dist_same_annulus.err | tail -5 | cut -c1-400
timeout 500 python tools/profile_parts.py --dofs 10000000 > gpurun_out/r2_profile_parts_10M.json 2> gpurun_out/r2_profile_parts_10M.err
echo "exit $?" >> gpurun_out/r2_profile_parts_10M.err
cat gpurun_out/r2_profile_parts_10M.json | cut -c1-3000; grep -v "Warn\|sparse_coo" gpurun_out/r2_profile_parts_10M.err | tail -3 | cut -c1-300
