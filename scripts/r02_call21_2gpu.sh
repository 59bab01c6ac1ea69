#!/usr/bin/env bash
# Round-2 two-GPU call on the build of record (wide backward tiles, 74-tile chunks): every 2-rank NCCL test (data-parallel
# step incl. ragged shards, vocab-parallel head, Ulysses glue, fused d_weight reduce-scatter in both transports - with a
# wide-tile case), then bench.py --gpus 2 with the NCCL all-reduce after the step and with the fused reduce-scatter.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "two_ranks" -p no:cacheprovider -rA > gpurun_out/r02u_two_rank_tests.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r02u_two_rank_tests.log
tail -12 gpurun_out/r02u_two_rank_tests.log
for mode in allreduce scatter; do
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29617 \
        bench.py --gpus 2 --steps 4 --warmup 3 --no-other --dw-reduce $mode > gpurun_out/r02u_bench_n2_${mode}.json 2> gpurun_out/r02u_bench_n2_${mode}.err
    echo "bench $mode rc=$?"; tail -2 gpurun_out/r02u_bench_n2_${mode}.err | cut -c1-300
    python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r02u_bench_n2_${mode}.json") if l.startswith("{")][-1])
print("$mode", {k: d[k] for k in ("value", "ms_per_step", "allreduce_exposed_ms", "phases_ms")}, d["dp_check"])
print({k: (round(v["ms_per_launch"], 3), round(v["frac"], 3)) for k, v in d["roofline_kernels"].items()})
PY
done
