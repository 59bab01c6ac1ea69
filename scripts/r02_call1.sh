#!/usr/bin/env bash
# Round-2 first GPU call: the whole GPU suite incl. the opt-in tests, the config 3/4/5 shard throughput, and a short
# A/B (base / stored-logits LSE / 74-tile chunk) on one box.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r02_call1_box.txt
SKYB200_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests -m gpu -q -n 4 -p no:cacheprovider \
    > gpurun_out/r02a_gpu_tests.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r02a_gpu_tests.log
tail -15 gpurun_out/r02a_gpu_tests.log
timeout 600 python scripts/bench_configs.py steps=2 warmup=1 > gpurun_out/r02a_bench_configs.jsonl 2> gpurun_out/r02a_bench_configs.err
echo "bench_configs rc=$?"; cat gpurun_out/r02a_bench_configs.jsonl; tail -3 gpurun_out/r02a_bench_configs.err
out=gpurun_out/r02a_ab.jsonl
: > "$out"
run() { local tag=$1; shift; env "$@" timeout 120 python scripts/step_ab.py tag="$tag" steps=3 warmup=2 >> "$out" 2>> gpurun_out/r02a_ab.err || echo "{\"tag\": \"$tag\", \"failed\": true}" >> "$out"; }
for round in 1 2; do
    run base A=1
    run stored_lse SKYB200_STORED_LSE=1
    run chunk74 SKYB200_CHUNK_TILES=74
done
cat "$out"
