#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
PYTHONPATH=$PWD timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 29541 tests/dp_worker.py > gpurun_out/r02g_dp_worker.log 2>&1
echo "dp_worker rc=$?"
grep -n "check failed\|DP_\|Error\|error\|Traceback" -A6 gpurun_out/r02g_dp_worker.log | head -60
