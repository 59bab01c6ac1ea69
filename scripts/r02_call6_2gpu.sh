#!/usr/bin/env bash
# Round-2 two-GPU call: every 2-rank NCCL test (data-parallel step incl. ragged shards, vocab-parallel head, Ulysses
# sequence-parallel glue), then bench.py at N = 2 (dp_check, all-reduce time with the rank skew taken out).
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02f_gpus.txt
timeout 600 python -m pytest tests -m gpu -q -k "two_ranks" -p no:cacheprovider -rA > gpurun_out/r02f_two_rank_tests.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/r02f_two_rank_tests.log
tail -12 gpurun_out/r02f_two_rank_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
    bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02f_bench_n2.json 2> gpurun_out/r02f_bench_n2.err
echo "bench rc=$?"; tail -3 gpurun_out/r02f_bench_n2.err | cut -c1-300
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02f_bench_n2.json") if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "tokens_per_s_per_gpu", "dp_check", "allreduce_exposed_ms", "phases_ms")})
print({k: (v.get("tokens_per_s"), v.get("step_roofline", {}).get("frac"), v.get("ms_per_step")) if isinstance(v, dict) else v
       for k, v in (d.get("other_configs") or {}).items()})
PY
