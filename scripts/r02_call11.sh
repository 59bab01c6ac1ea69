#!/usr/bin/env bash
# Round-2 call 11: split-phase lockstep (SKYB200_LOCK_SLACK) and the new d_weight defaults (two N-groups of 7,
# evict-first A / C traffic) - isolated GEMMs first, then the whole step, alternating processes on one box.
set -u
mkdir -p gpurun_out
out=gpurun_out/r02k_bwd_ab.jsonl
: > "$out"
run() { local tag=$1; shift; env "$@" timeout 120 python scripts/bwd_gemm_probe.py quick=1 tag="$tag" >> "$out" 2>> gpurun_out/r02k.err || echo "{\"tag\": \"$tag\", \"failed\": true}" >> "$out"; }
OLD="SKYB200_DW_GROUP_N=14 SKYB200_HINT_DW_A=1 SKYB200_HINT_DW_C=1"
for round in 1 2; do
    run old $OLD
    run new X=1
    run new_slack8 SKYB200_LOCK_SLACK=8
    run new_slack16 SKYB200_LOCK_SLACK=16
    run new_slack32 SKYB200_LOCK_SLACK=32
    run new_lock64_slack16 SKYB200_DW_LOCKSTEP=64 SKYB200_DH_LOCKSTEP=64 SKYB200_LOCK_SLACK=16
    run new_lock256_slack32 SKYB200_DW_LOCKSTEP=256 SKYB200_DH_LOCKSTEP=256 SKYB200_LOCK_SLACK=32
    run new_b1 SKYB200_HINT_DW_B=1
done
cat "$out"
out=gpurun_out/r02k_step_ab.jsonl
: > "$out"
srun() { local tag=$1; shift; env "$@" timeout 150 python scripts/step_ab.py tag="$tag" steps=3 warmup=2 >> "$out" 2>> gpurun_out/r02k.err || echo "{\"tag\": \"$tag\", \"failed\": true}" >> "$out"; }
for round in 1 2; do
    srun old $OLD
    srun new X=1
    srun new_slack16 SKYB200_LOCK_SLACK=16
    srun new_slack32 SKYB200_LOCK_SLACK=32
    srun new_fwdlock_slack16 SKYB200_FWD_LOCKSTEP=128 SKYB200_LOCK_SLACK=16
done
cat "$out"
tail -3 gpurun_out/r02k.err
