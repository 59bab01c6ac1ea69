#!/usr/bin/env bash
# Round-2 two-GPU call: the peers' pieces of the fused d_weight reduce-scatter through TMA reduce-add on a tensor map
# over the PEER-MAPPED shard (instead of per-thread red.v4) - parity first (red mode, then tma mode), then N = 2 bench.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "fused_dweight_reduce_scatter" -p no:cacheprovider -rA > gpurun_out/r02l_dws_tests.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r02l_dws_tests.log
tail -12 gpurun_out/r02l_dws_tests.log
nvidia-smi --query-gpu=index,name,memory.used --format=csv
if grep -q "2 passed" gpurun_out/r02l_dws_tests.log; then
  for mode in scatter allreduce scatter; do
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29617 \
          bench.py --gpus 2 --steps 3 --warmup 3 --no-other --dw-reduce $mode > gpurun_out/r02l_bench_n2_${mode}.json 2> gpurun_out/r02l_bench_n2_${mode}.err
      echo "bench $mode rc=$?"; tail -2 gpurun_out/r02l_bench_n2_${mode}.err | cut -c1-300
      python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r02l_bench_n2_${mode}.json") if l.startswith("{")][-1])
print("$mode", {k: d[k] for k in ("value", "ms_per_step", "allreduce_exposed_ms", "phases_ms")}, d["dp_check"])
print({k: (round(v["ms_per_launch"], 3), round(v["frac"], 3)) for k, v in d["roofline_kernels"].items()})
PY
  done
fi
