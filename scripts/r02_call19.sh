#!/usr/bin/env bash
# Round-2 call 19: knob sweep for the wide backward GEMMs (lockstep period, L2 hints, d_weight N-groups) and the
# 148-tile chunk, alternating processes.
set -u
mkdir -p gpurun_out
out=gpurun_out/r02s_bwd_ab.jsonl
: > "$out"
run() { local tag=$1; shift; env "$@" timeout 150 python scripts/bwd_gemm_probe.py 18944 152064 3584 quick=1 tag="$tag" >> "$out" 2>> gpurun_out/r02s.err || echo "{\"tag\": \"$tag\", \"failed\": true}" >> "$out"; }
run base X=1
run lock64 SKYB200_DW_LOCKSTEP=64 SKYB200_DH_LOCKSTEP=64
run lock296 SKYB200_DW_LOCKSTEP=296 SKYB200_DH_LOCKSTEP=297
run lock32 SKYB200_DW_LOCKSTEP=32 SKYB200_DH_LOCKSTEP=32
run dw_a1 SKYB200_HINT_DW_A=1
run dw_c1 SKYB200_HINT_DW_C=1
run dw_n4 SKYB200_DW_GROUP_N=4
run dh_b3 SKYB200_HINT_DH_B=3
run dh_a2 SKYB200_HINT_DH_A=2
run base X=1
cat "$out" | cut -c1-260
out=gpurun_out/r02s_step_ab.jsonl
: > "$out"
srun() { local tag=$1; shift; env "$@" timeout 170 python scripts/step_ab.py tag="$tag" steps=3 warmup=2 >> "$out" 2>> gpurun_out/r02s.err || echo "{\"tag\": \"$tag\", \"failed\": true}" >> "$out"; }
srun base X=1
srun chunk148 SKYB200_CHUNK_TILES=148
srun base X=1
cat "$out"; tail -5 gpurun_out/r02s.err
