#!/usr/bin/env bash
# Round-2 call 13: per-kernel ncu view of the memory-bound kernels (where do the 31.8 us of the loss call go?), the
# same kernels on 4x the batch, and the evict-first store of the d(logits) transform pass inside the step.
set -u
mkdir -p gpurun_out
python scripts/hbm_probe.py > gpurun_out/r02m_hbm_probe.log 2> gpurun_out/r02m.err
echo "hbm rc=$?"; cat gpurun_out/r02m_hbm_probe.log; tail -3 gpurun_out/r02m.err
python scripts/hbm_probe.py quick > gpurun_out/r02m_hbm_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__inst_executed_pipe_xu.sum \
    --clock-control none -k regex:"loss_|group_|row_sum|gae_|whiten|reward_kl|row_means" -c 400 --csv \
    --log-file gpurun_out/r02m_hbm_kernels_ncu.csv python scripts/hbm_probe.py quick > gpurun_out/r02m_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r02m_ncu.log | cut -c1-200
out=gpurun_out/r02m_step_ab.jsonl
: > "$out"
srun() { local tag=$1; shift; env "$@" timeout 150 python scripts/step_ab.py tag="$tag" steps=3 warmup=2 >> "$out" 2>> gpurun_out/r02m.err || echo "{\"tag\": \"$tag\", \"failed\": true}" >> "$out"; }
for round in 1 2; do
    srun base X=1
    srun tr_evict SKYB200_TRANSFORM_EVICT=1
done
cat "$out"
