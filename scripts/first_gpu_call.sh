#!/usr/bin/env bash
# Everything that was written after the round-1 GPU budget ran out, in one gpurun call (DESIGN.md "What comes next",
# item 0).  Logs land in gpurun_out/.
#
#   gpurun --timeout 1500 -- 'bash scripts/first_gpu_call.sh'
set -u
mkdir -p gpurun_out
# 1. the whole GPU parity suite, including the opt-in test of the default-off SKYB200_STORED_LSE kernels
SKYB200_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests -m gpu -q -s -p no:cacheprovider \
    > gpurun_out/gpu_tests.log 2>&1
echo "pytest rc=$?" | tee -a gpurun_out/gpu_tests.log
tail -5 gpurun_out/gpu_tests.log
# 2. the fused step on the config 3 / 4 / 5 shards
timeout 600 python scripts/bench_configs.py steps=3 warmup=2 > gpurun_out/bench_configs.jsonl 2> gpurun_out/bench_configs.err
echo "bench_configs rc=$?"; cat gpurun_out/bench_configs.jsonl
# 3. A/B of the environment knobs (chunk size, lockstep period, stored-logits lse) on this box
bash scripts/ab_matrix.sh > /dev/null
python - <<'PY'
import json, collections
runs = collections.defaultdict(list)
for line in open("gpurun_out/ab_matrix.jsonl"):
    try:
        d = json.loads(line)
    except ValueError:
        continue
    if "ms_per_step" in d:
        runs[d["tag"]].append(d["ms_per_step"])
for tag, ms in runs.items():
    print(f"{tag:20s} " + "  ".join(f"{x:8.1f}" for x in ms) + f"   mean {sum(ms) / len(ms):8.1f} ms")
PY
