#!/usr/bin/env bash
# A/B matrix for the fused step on ONE box (box-to-box variance is larger than most effects, so variants alternate
# inside one gpurun call).  Every line of gpurun_out/ab_matrix.jsonl is one scripts/step_ab.py run.
#
#   gpurun --timeout 900 -- 'bash scripts/ab_matrix.sh'
#
# Variants (environment knobs of libskyb200, see README): the token chunk (37 vs 74 pair-tiles: the fp32 d_weight
# read-modify-write and the weight re-read per chunk halve, the bf16 logits store doubles to 2 x 5.8 GB), the
# lockstep period of the deep-K GEMMs, lockstep on the forward GEMM, and the forward tile group.
set -u
mkdir -p gpurun_out
out=gpurun_out/ab_matrix.jsonl
: > "$out"
run() {  # run <tag> <extra step_ab.py argument or -> [ENV=VALUE ...]
    local tag=$1 extra=$2; shift 2
    [ "$extra" = "-" ] && extra=""
    env "$@" timeout 120 python scripts/step_ab.py tag="$tag" steps=3 warmup=2 $extra >> "$out" 2>> gpurun_out/ab_matrix.err \
        || echo "{\"tag\": \"$tag\", \"failed\": true}" >> "$out"
}
for round in 1 2; do
    run base -
    run chunk74 - SKYB200_CHUNK_TILES=74
    run lock64 - SKYB200_DH_LOCKSTEP=64 SKYB200_DW_LOCKSTEP=64
    run lock256 - SKYB200_DH_LOCKSTEP=256 SKYB200_DW_LOCKSTEP=256
    run fwdlock128 - SKYB200_FWD_LOCKSTEP=128
    run chunk74_fwdgrp37 - SKYB200_CHUNK_TILES=74 SKYB200_FWD_GROUP_M=37
    run chunk148 - SKYB200_CHUNK_TILES=148
    run stored_lse - SKYB200_STORED_LSE=1
    run ragged_base ragged=1
done
cat "$out"
