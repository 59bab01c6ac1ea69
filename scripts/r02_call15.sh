#!/usr/bin/env bash
# Round-2 call 15: the numbers of record for this build - the default bench line exactly as the driver runs it, the
# reference arm, `ncu --set full` of the three in-step GEMMs (-> profiles/r02_traffic.json), the launch list of one
# full-size step and the per-kernel view of the memory-bound kernels.
set -u
mkdir -p gpurun_out
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02o_bench_n1.json 2> gpurun_out/r02o_bench_n1.err
echo "bench rc=$?"; tail -3 gpurun_out/r02o_bench_n1.err | cut -c1-300; cut -c1-400 gpurun_out/r02o_bench_n1.json
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r02o_bench_ref.json 2> gpurun_out/r02o_bench_ref.err
echo "ref rc=$?"; cut -c1-300 gpurun_out/r02o_bench_ref.json
CMD="python bench.py --workload c2s --steps 1 --warmup 1 --no-other"
$CMD > gpurun_out/r02o_plain_c2s.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:gemm_sm100_kernel -s 16 -c 12 \
    -f -o gpurun_out/r02o_gemm_full $CMD > gpurun_out/r02o_ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/r02o_ncu_full.log | cut -c1-200
CMD2="python bench.py --steps 1 --warmup 1 --no-other"
$CMD2 > gpurun_out/r02o_plain_c2.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv \
    --log-file gpurun_out/r02o_launches_c2.csv $CMD2 > gpurun_out/r02o_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
python scripts/hbm_probe.py quick > gpurun_out/r02o_hbm_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum \
    --clock-control none -k regex:"loss_|group_|row_sum|gae_|whiten|reward_kl|row_means" -c 400 --csv \
    --log-file gpurun_out/r02o_hbm_kernels_ncu.csv python scripts/hbm_probe.py quick > gpurun_out/r02o_ncu_hbm.log 2>&1
echo "ncu hbm rc=$?"
ls -la gpurun_out | grep r02o
