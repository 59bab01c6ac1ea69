#!/usr/bin/env bash
# Round-2 third GPU call: TMA-store forward epilogue - parity, epilogue microbench, step A/B vs round 1, new bench.py.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -n 4 -p no:cacheprovider > gpurun_out/r02c_gpu_tests.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r02c_gpu_tests.log
tail -8 gpurun_out/r02c_gpu_tests.log
timeout 200 python scripts/gemm_tune.py 9472 152064 3584 > gpurun_out/r02c_gemm_tune.jsonl 2> gpurun_out/r02c.err
timeout 200 python scripts/gemm_tune.py 9472 151936 896 >> gpurun_out/r02c_gemm_tune.jsonl 2>> gpurun_out/r02c.err
cat gpurun_out/r02c_gemm_tune.jsonl
out=gpurun_out/r02c_ab.jsonl
: > "$out"
run() { local tag=$1; shift; env "$@" timeout 120 python scripts/step_ab.py tag="$tag" steps=3 warmup=2 >> "$out" 2>> gpurun_out/r02c.err || echo "{\"tag\": \"$tag\", \"failed\": true}" >> "$out"; }
for round in 1 2; do
    (cd scratch/r01_tree && timeout 120 python scripts/step_ab.py tag=r01 steps=3 warmup=2) >> "$out" 2>> gpurun_out/r02c.err
    run new A=1
    run new_chunk74 SKYB200_CHUNK_TILES=74
done
cat "$out"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r02c_bench.err; cat gpurun_out/r02c_bench.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02c_bench_ref.json 2>> gpurun_out/r02c_bench.err
cat gpurun_out/r02c_bench_ref.json
