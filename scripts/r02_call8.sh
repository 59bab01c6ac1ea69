#!/usr/bin/env bash
# Round-2 call 8 (after the container re-creation): validate the committed build end to end - the whole GPU suite, the
# default bench line as the driver runs it, `ncu --set full` of the three in-step GEMMs (-> profiles/r02_traffic.json)
# and the launch list of one full-size step.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -n 4 -p no:cacheprovider -rs > gpurun_out/r02h_gpu_tests.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r02h_gpu_tests.log
tail -15 gpurun_out/r02h_gpu_tests.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r02h_bench.err | cut -c1-300; cut -c1-600 gpurun_out/r02h_bench.json
# ncu --set full: the 8-sequence slice of config 2 (4 chunks per step, 3 of them full); setup forward = 4 launches,
# warm-up step = 12, timed step = 12 GEMM launches
CMD="python bench.py --workload c2s --steps 1 --warmup 1 --no-other"
$CMD > gpurun_out/r02h_plain_c2s.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:gemm_sm100_kernel -s 16 -c 12 \
    -f -o gpurun_out/r02h_gemm_full $CMD > gpurun_out/r02h_ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/r02h_ncu_full.log | cut -c1-300
CMD2="python bench.py --steps 1 --warmup 1 --no-other"
$CMD2 > gpurun_out/r02h_plain_c2.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/r02h_launches_c2.csv $CMD2 > gpurun_out/r02h_ncu_launches.log 2>&1
echo "ncu launches rc=$?"; tail -2 gpurun_out/r02h_ncu_launches.log | cut -c1-300
ls -la gpurun_out | head -30
