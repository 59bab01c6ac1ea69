#!/usr/bin/env bash
# Round-2 fifth GPU call: d_weight epilogue through TMA reduce-add, GAE rows split over blocks, fast exp in the loss pass.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -n 4 -p no:cacheprovider > gpurun_out/r02e_gpu_tests.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r02e_gpu_tests.log
tail -8 gpurun_out/r02e_gpu_tests.log
timeout 600 python scripts/bwd_gemm_probe.py > gpurun_out/r02e_bwd_probe.jsonl 2> gpurun_out/r02e.err
cat gpurun_out/r02e_bwd_probe.jsonl; tail -3 gpurun_out/r02e.err
timeout 300 python scripts/hbm_probe.py > gpurun_out/r02e_hbm_probe.log 2>> gpurun_out/r02e.err
cat gpurun_out/r02e_hbm_probe.log
out=gpurun_out/r02e_ab.jsonl
: > "$out"
run() { local tag=$1; shift; env "$@" timeout 150 python scripts/step_ab.py tag="$tag" steps=3 warmup=2 >> "$out" 2>> gpurun_out/r02e.err || echo "{\"tag\": \"$tag\", \"failed\": true}" >> "$out"; }
for round in 1 2; do
    run new A=1
    run fwdlock128 SKYB200_FWD_LOCKSTEP=128
    run fwdlock64 SKYB200_FWD_LOCKSTEP=64
    run fwdlock256 SKYB200_FWD_LOCKSTEP=256
done
cat "$out"
