#!/usr/bin/env bash
# Round-2 two-GPU call: the d_weight GEMM whose epilogue reduce-scatters the gradient over NVLink peer memory -
# parity against NCCL's reduce-scatter (tests/dws_worker.py), then bench.py at N = 2 in both reduction modes,
# alternating (box-to-box and run-to-run variance is larger than the effect).
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02j_topo.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q -k "two_ranks" -p no:cacheprovider -rA > gpurun_out/r02j_two_rank_tests.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r02j_two_rank_tests.log
tail -12 gpurun_out/r02j_two_rank_tests.log
if grep -q "DWS_OK\|4 passed" gpurun_out/r02j_two_rank_tests.log; then
  for round in 1 2; do
    for mode in allreduce scatter; do
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2961$round \
          bench.py --gpus 2 --steps 3 --warmup 3 --no-other --dw-reduce $mode > gpurun_out/r02j_bench_n2_${mode}_$round.json 2> gpurun_out/r02j_bench_n2_${mode}_$round.err
      echo "bench $mode $round rc=$?"; tail -2 gpurun_out/r02j_bench_n2_${mode}_$round.err | cut -c1-300
      python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/r02j_bench_n2_${mode}_$round.json") if l.startswith("{")][-1])
print("$mode", {k: d[k] for k in ("value", "ms_per_step", "allreduce_exposed_ms", "phases_ms")}, d["dp_check"])
print({k: (round(v["ms_per_launch"], 3), round(v["frac"], 3)) for k, v in d["roofline_kernels"].items()})
PY
    done
  done
fi
