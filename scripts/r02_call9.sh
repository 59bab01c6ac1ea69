#!/usr/bin/env bash
# Round-2 call 9: (1) parity of the new one-launch GRPO cluster kernel and the restructured loss token pass,
# (2) their HBM numbers, (3) A/B of the L2 policy / tile-order knobs of the two backward GEMMs (ncu r02h: d_weight is
# only 80 % tensor-active and re-reads the 68 MB hidden chunk from DRAM once per round of M tiles).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_advantage_gpu.py tests/test_loss_gpu.py tests/test_policy_update_gpu.py tests/test_zproperty_gpu.py \
    -m gpu -q -n 4 -p no:cacheprovider -x > gpurun_out/r02i_gpu_tests.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r02i_gpu_tests.log
tail -15 gpurun_out/r02i_gpu_tests.log
timeout 300 python scripts/hbm_probe.py > gpurun_out/r02i_hbm_probe.log 2> gpurun_out/r02i.err
cat gpurun_out/r02i_hbm_probe.log; tail -3 gpurun_out/r02i.err
out=gpurun_out/r02i_bwd_ab.jsonl
: > "$out"
run() { local tag=$1; shift; env "$@" timeout 120 python scripts/bwd_gemm_probe.py quick=1 tag="$tag" >> "$out" 2>> gpurun_out/r02i.err || echo "{\"tag\": \"$tag\", \"failed\": true}" >> "$out"; }
for round in 1 2; do
    run base X=1
    run dw_n7 SKYB200_DW_GROUP_N=7
    run dw_a2 SKYB200_HINT_DW_A=2
    run dw_c2 SKYB200_HINT_DW_C=2
    run dw_a2c2 SKYB200_HINT_DW_A=2 SKYB200_HINT_DW_C=2
    run dw_n7a2c2 SKYB200_DW_GROUP_N=7 SKYB200_HINT_DW_A=2 SKYB200_HINT_DW_C=2
    run dw_n5a2c2 SKYB200_DW_GROUP_N=5 SKYB200_HINT_DW_A=2 SKYB200_HINT_DW_C=2
    run dh_a2b2 SKYB200_HINT_DH_A=2 SKYB200_HINT_DH_B=2
    run dh_g5 SKYB200_DH_GROUP_M=5
done
cat "$out"
