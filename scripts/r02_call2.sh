#!/usr/bin/env bash
# Round-2 second GPU call: parity of the slab-relative tile store + off-policy correction in the fused step, then the
# "what-if" A/B (which part of the step is worth removing) against the round-1 tree on the same box.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -n 4 -p no:cacheprovider > gpurun_out/r02b_gpu_tests.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r02b_gpu_tests.log
tail -8 gpurun_out/r02b_gpu_tests.log
timeout 120 python -m pytest tests/test_zconfigs_gpu.py -m gpu -q -s -k "logit_scale or peaked" -p no:cacheprovider 2>&1 | grep "^std" 
out=gpurun_out/r02b_ab.jsonl
: > "$out"
run() { local tag=$1; shift; env "$@" timeout 120 python scripts/step_ab.py tag="$tag" steps=3 warmup=2 >> "$out" 2>> gpurun_out/r02b_ab.err || echo "{\"tag\": \"$tag\", \"failed\": true}" >> "$out"; }
for round in 1 2; do
    (cd scratch/r01_tree && timeout 120 python scripts/step_ab.py tag=r01 steps=3 warmup=2) >> "$out" 2>> gpurun_out/r02b_ab.err
    run new A=1
    run no_transform SKYB200_WHATIF=1
    run no_rmw SKYB200_WHATIF=2
    run no_store SKYB200_WHATIF=4
    run no_transform_rmw SKYB200_WHATIF=3
    run new_chunk74 SKYB200_CHUNK_TILES=74
    run new_chunk148 SKYB200_CHUNK_TILES=148
done
cat "$out"
timeout 200 python scripts/gemm_tune.py 9472 152064 3584 > gpurun_out/r02b_gemm_tune.jsonl 2>> gpurun_out/r02b_ab.err
timeout 200 python scripts/gemm_tune.py 9472 151936 896 >> gpurun_out/r02b_gemm_tune.jsonl 2>> gpurun_out/r02b_ab.err
(cd scratch/r01_tree && timeout 200 python scripts/gemm_tune.py 9472 152064 3584 && timeout 200 python scripts/gemm_tune.py 9472 151936 896) > gpurun_out/r02b_gemm_tune_r01.jsonl 2>> gpurun_out/r02b_ab.err
cat gpurun_out/r02b_gemm_tune.jsonl gpurun_out/r02b_gemm_tune_r01.jsonl
