#!/usr/bin/env bash
# Round-2 call 18: the wide backward tiles as the default - whole GPU suite, d_hidden group sweep, default bench line
# and the reference arm.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -n 4 -p no:cacheprovider -rs > gpurun_out/r02r_gpu_tests.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r02r_gpu_tests.log
tail -12 gpurun_out/r02r_gpu_tests.log
out=gpurun_out/r02r_bwd_ab.jsonl
: > "$out"
run() { local tag=$1 rows=$2; shift 2; env "$@" timeout 150 python scripts/bwd_gemm_probe.py "$rows" 152064 3584 quick=1 tag="$tag" >> "$out" 2>> gpurun_out/r02r.err || echo "{\"tag\": \"$tag\", \"failed\": true}" >> "$out"; }
run g11 18944 X=1
run g7 18944 SKYB200_DH_GROUP_M_WIDE=7
run g5 18944 SKYB200_DH_GROUP_M_WIDE=5
run g15 18944 SKYB200_DH_GROUP_M_WIDE=15
run g11 18944 X=1
run g7 18944 SKYB200_DH_GROUP_M_WIDE=7
cat "$out"
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02r_bench_n1.json 2> gpurun_out/r02r_bench_n1.err
echo "bench rc=$?"; tail -3 gpurun_out/r02r_bench_n1.err | cut -c1-300; cut -c1-600 gpurun_out/r02r_bench_n1.json
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r02r_bench_ref.json 2> gpurun_out/r02r_bench_ref.err
echo "ref rc=$?"; cut -c1-300 gpurun_out/r02r_bench_ref.json
