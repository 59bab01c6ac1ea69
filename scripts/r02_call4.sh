#!/usr/bin/env bash
# Round-2 fourth GPU call: where do the backward GEMMs lose time inside the step?  kernel form vs operand data vs
# co-running auxiliary kernels.
set -u
mkdir -p gpurun_out
timeout 600 python scripts/bwd_gemm_probe.py > gpurun_out/r02d_bwd_probe.jsonl 2> gpurun_out/r02d.err
cat gpurun_out/r02d_bwd_probe.jsonl; tail -3 gpurun_out/r02d.err
out=gpurun_out/r02d_ab.jsonl
: > "$out"
run() { local tag=$1; shift; env "$@" timeout 150 python scripts/step_ab.py tag="$tag" steps=3 warmup=2 >> "$out" 2>> gpurun_out/r02d.err || echo "{\"tag\": \"$tag\", \"failed\": true}" >> "$out"; }
run_t0() { timeout 150 python scripts/step_ab.py tag="$1" steps=3 warmup=2 timing=0 >> "$out" 2>> gpurun_out/r02d.err; }
for round in 1 2; do
    run new A=1
    run aux_serial SKYB200_AUX_SERIAL=1
    run_t0 no_timing_events
    run chunk74 SKYB200_CHUNK_TILES=74
    run lock64 SKYB200_DH_LOCKSTEP=64 SKYB200_DW_LOCKSTEP=64
    run lock0 SKYB200_DH_LOCKSTEP=0 SKYB200_DW_LOCKSTEP=0
    run fwdlock128 SKYB200_FWD_LOCKSTEP=128
done
cat "$out"
