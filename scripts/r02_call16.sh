#!/usr/bin/env bash
# Round-2 call 16: L2 prefetch of the TMA producer (cp.async.bulk.prefetch.tensor, P K blocks ahead of the load) for
# the two backward GEMMs - isolated sustained rates, then the whole step.
set -u
mkdir -p gpurun_out
out=gpurun_out/r02p_bwd_ab.jsonl
: > "$out"
run() { local tag=$1; shift; env "$@" timeout 120 python scripts/bwd_gemm_probe.py quick=1 tag="$tag" >> "$out" 2>> gpurun_out/r02p.err || echo "{\"tag\": \"$tag\", \"failed\": true}" >> "$out"; }
for round in 1 2; do
    run base X=1
    run pf4 SKYB200_DH_PREFETCH=4 SKYB200_DW_PREFETCH=4
    run pf8 SKYB200_DH_PREFETCH=8 SKYB200_DW_PREFETCH=8
    run pf16 SKYB200_DH_PREFETCH=16 SKYB200_DW_PREFETCH=16
    run pf32 SKYB200_DH_PREFETCH=32 SKYB200_DW_PREFETCH=32
done
cat "$out"
out=gpurun_out/r02p_step_ab.jsonl
: > "$out"
srun() { local tag=$1; shift; env "$@" timeout 150 python scripts/step_ab.py tag="$tag" steps=3 warmup=2 >> "$out" 2>> gpurun_out/r02p.err || echo "{\"tag\": \"$tag\", \"failed\": true}" >> "$out"; }
for round in 1 2; do
    srun base X=1
    srun pf8 SKYB200_DH_PREFETCH=8 SKYB200_DW_PREFETCH=8
    srun pf16 SKYB200_DH_PREFETCH=16 SKYB200_DW_PREFETCH=16
    srun pf8_all SKYB200_DH_PREFETCH=8 SKYB200_DW_PREFETCH=8 SKYB200_FWD_PREFETCH=8
done
cat "$out"; tail -3 gpurun_out/r02p.err
