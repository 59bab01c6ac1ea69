#!/usr/bin/env bash
# Round-2 call 20: ncu evidence for the build of record (wide backward tiles): `--set full` of the three in-step GEMMs
# on the 16-sequence slice of config 2 (-> profiles/r02_traffic.json) and the launch list of one full-size step.
set -u
mkdir -p gpurun_out
CMD="python bench.py --workload c2s --steps 1 --warmup 1 --no-other"
$CMD > gpurun_out/r02t_plain_c2s.log 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:gemm_sm100_kernel -s 16 -c 12 \
    -f -o gpurun_out/r02t_gemm_full $CMD > gpurun_out/r02t_ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/r02t_ncu_full.log | cut -c1-200
CMD2="python bench.py --steps 1 --warmup 1 --no-other"
$CMD2 > gpurun_out/r02t_plain_c2.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv \
    --log-file gpurun_out/r02t_launches_c2.csv $CMD2 > gpurun_out/r02t_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
ls -la gpurun_out | grep r02t
