#!/usr/bin/env bash
# Round-2 call 22: forward tile-order group with the 74-tile chunks (19 was tuned for 37-tile chunks; the forward now
# reads 13.5 GB of DRAM per launch, 3x its ideal) - whole step, alternating processes.
set -u
mkdir -p gpurun_out
out=gpurun_out/r02v_step_ab.jsonl
: > "$out"
srun() { local tag=$1; shift; env "$@" timeout 170 python scripts/step_ab.py tag="$tag" steps=3 warmup=2 >> "$out" 2>> gpurun_out/r02v.err || echo "{\"tag\": \"$tag\", \"failed\": true}" >> "$out"; }
srun g19 X=1
srun g25 SKYB200_FWD_GROUP_M=25
srun g37 SKYB200_FWD_GROUP_M=37
srun g13 SKYB200_FWD_GROUP_M=13
srun g19 X=1
srun g19_lock64 SKYB200_FWD_LOCKSTEP=64
cat "$out"; tail -5 gpurun_out/r02v.err
