"""bench.py contract on CPU: the reference arm (the oracle port timed on host cores) prints ONE JSON line with the
keys the driver reads; the two arms describe the same workload with the same `config` object."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=280, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-1500:]
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1
    assert d["unit"] == "tokens/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_both_arms_share_metric_and_config():
    sys.path.insert(0, ROOT)
    import bench

    w = bench.WORKLOADS["c2"]
    cfg = bench.config_dict(w, 1)
    assert "Qwen2.5-7B head" in cfg["workload"] and "256 seqs x 4096 tokens" in cfg["workload"]
    assert cfg["per_gpu_tokens"] == 256 * 4096
    assert bench.METRIC.startswith("policy-update tokens/sec/GPU")
    peaks = bench.measured_peaks()
    assert peaks["tflops"] > 0 and peaks["hbm"] > 0
