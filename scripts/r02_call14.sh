#!/usr/bin/env bash
# Round-2 call 14: validation of the build (whole GPU suite), the memory-bound kernels with the new timing method,
# and the other configs after the d_weight group-size model (c4-gspo: one 32768-token sequence per chunk).
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -n 4 -p no:cacheprovider -rs > gpurun_out/r02n_gpu_tests.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r02n_gpu_tests.log
tail -12 gpurun_out/r02n_gpu_tests.log
timeout 300 python -m pytest tests/test_zconfigs_gpu.py -m gpu -q -s -k "logit_scale or peaked" -p no:cacheprovider 2>&1 | grep -E "^std|passed|failed" | tee gpurun_out/r02n_backward_error.log
timeout 300 python scripts/hbm_probe.py > gpurun_out/r02n_hbm_probe.log 2> gpurun_out/r02n.err
cat gpurun_out/r02n_hbm_probe.log | head -16; tail -3 gpurun_out/r02n.err
timeout 900 python scripts/bench_configs.py steps=2 warmup=1 > gpurun_out/r02n_bench_configs.jsonl 2>> gpurun_out/r02n.err
echo "bench_configs rc=$?"; cut -c1-420 gpurun_out/r02n_bench_configs.jsonl
