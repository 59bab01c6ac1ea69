#!/usr/bin/env bash
# Round-2 call 17: wide (256 x 512) tiles for the two deep-K backward GEMMs (SKYB200_WIDE: 1 = d_hidden, 2 = d_weight)
# with 74-tile token chunks - correctness, isolated sustained rates, then the whole step, alternating processes.
set -u
mkdir -p gpurun_out
log=gpurun_out/r02q_tests.log
timeout 400 python -m pytest tests/test_head_gpu.py tests/test_policy_update_gpu.py -m gpu -q -k "gemm_mainloop or wide" \
    -p no:cacheprovider > "$log" 2>&1
echo "pytest rc=$?" | tee -a "$log"; tail -15 "$log"
out=gpurun_out/r02q_bwd_ab.jsonl
: > "$out"
run() { local tag=$1 rows=$2; shift 2; env "$@" timeout 150 python scripts/bwd_gemm_probe.py "$rows" 152064 3584 quick=1 tag="$tag" >> "$out" 2>> gpurun_out/r02q.err || echo "{\"tag\": \"$tag\", \"failed\": true}" >> "$out"; }
run base_9472 9472 X=1
run base_18944 18944 X=1
run wide_18944 18944 SKYB200_WIDE=3
run wide_9472_forced 9472 SKYB200_WIDE=7
run base_18944 18944 X=1
run wide_18944 18944 SKYB200_WIDE=3
run wide_18944_g7 18944 SKYB200_WIDE=3 SKYB200_DH_GROUP_M_WIDE=7
run wide_18944_nolock 18944 SKYB200_WIDE=3 SKYB200_DH_LOCKSTEP=0 SKYB200_DW_LOCKSTEP=0
cat "$out"
out=gpurun_out/r02q_step_ab.jsonl
: > "$out"
srun() { local tag=$1; shift; env "$@" timeout 170 python scripts/step_ab.py tag="$tag" steps=3 warmup=2 >> "$out" 2>> gpurun_out/r02q.err || echo "{\"tag\": \"$tag\", \"failed\": true}" >> "$out"; }
srun base X=1
srun wide3 SKYB200_WIDE=3
srun chunk74 SKYB200_CHUNK_TILES=74
srun wide1 SKYB200_WIDE=1
srun wide2 SKYB200_WIDE=2
srun base X=1
srun wide3 SKYB200_WIDE=3
cat "$out"; tail -5 gpurun_out/r02q.err
