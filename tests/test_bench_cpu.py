"""bench.py contract on CPU: the reference arm (the UNMODIFIED reference from /root/reference or baseline/_ref timed on
host cores; the oracle port only when neither exists) prints ONE JSON line with the keys the driver reads; the two arms
describe the same workload with the same `config` object; the traffic table is tied to the kernel sources."""
import json
import os
import subprocess
import sys

import pytest

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
    sys.path.insert(0, ROOT)
    from oracle import ref_loader

    assert d["cpu_baseline"]["kind"] == ("reference" if ref_loader.available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1
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


def test_reference_arm_falls_back_to_the_port_without_a_copy_of_the_reference(tmp_path):
    env = dict(os.environ, SKYRL_REF_ROOT=str(tmp_path / "nowhere"))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=280, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-1500:]
    d = json.loads([ln for ln in r.stdout.strip().splitlines() if ln.startswith("{")][0])
    assert d["cpu_baseline"]["kind"] == "port"


def test_vendored_reference_is_byte_identical_and_drives_the_step():
    """__graft_entry__.build() copies the four hot-path files of the reference, unmodified, to baseline/_ref/ (the GPU
    box has no /root/reference); the baseline step runs from that copy and matches the oracle."""
    sys.path.insert(0, ROOT)
    from baseline import ref_path

    if not os.path.isdir(ref_path.SOURCE_DIR):
        pytest.skip("/root/reference not present")
    assert ref_path.vendor()
    for rel in ref_path.FILES:
        assert open(os.path.join(ref_path.SOURCE_DIR, rel), "rb").read() == \
            open(os.path.join(ref_path.VENDOR_DIR, rel), "rb").read(), rel
    env = dict(os.environ, SKYRL_REF_ROOT=ref_path.VENDOR_DIR)
    code = ("import sys; sys.path.insert(0, %r); import bench; from baseline import ref_path; "
            "print(ref_path.cpu_step(bench.WORKLOADS['tiny'], bench.make_cfg(), 512, 2)[2])" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=280, env=env)
    assert r.returncode == 0 and r.stdout.strip().endswith("reference"), r.stdout[-500:] + r.stderr[-1500:]


def test_traffic_table_is_refused_when_the_kernel_sources_changed(tmp_path, monkeypatch):
    import bench

    good = {"source_hash": bench.kernel_source_hash(), "H": 3584, "V": 152064, "capture": "unit test",
            "kernels": {"forward": {"dram_read_bytes": 1.0, "dram_write_bytes": 2.0}}}
    prof = tmp_path / "profiles"
    prof.mkdir()
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    monkeypatch.setattr(bench, "kernel_source_hash", lambda: good["source_hash"])
    (prof / "r02_traffic.json").write_text(json.dumps(good))
    assert bench.measured_traffic(3584, 152064)[0] == {"forward": 3.0}
    assert bench.measured_traffic(896, 151936)[0] == {}                      # another head shape
    (prof / "r02_traffic.json").write_text(json.dumps(dict(good, source_hash="0" * 16)))
    table, note = bench.measured_traffic(3584, 152064)
    assert table == {} and note.startswith("STALE")                          # loud, and traffic = null
    (prof / "r02_traffic.json").unlink()
    assert bench.measured_traffic(3584, 152064)[0] == {}


def test_clock_sampler_parses_nvidia_smi_lines():
    """The `clocks` object of the bench line (median SM clock under load, max clock, throttle reasons) from the
    nvidia-smi CSV the sampler reads (B200_PROFILING.md recipe); malformed lines are skipped, not fatal."""
    import bench

    class _Done:
        def terminate(self):
            pass

        def wait(self, timeout=None):
            return 0

        def kill(self):
            pass

    s = bench.ClockSampler(0)
    s.proc = _Done()
    s.lines = [
        "0, 1965, 1965, 180.2, 0x0000000000000000, Not Active, Not Active, Not Active, Not Active",
        "0, 1320, 1965, 995.1, 0x0000000000000004, Not Active, Not Active, Not Active, Active",
        "0, 1305, 1965, 998.7, 0x0000000000000004, Not Active, Not Active, Not Active, Active",
        "0, [N/A], 1965, [N/A], 0x0, Not Active, Not Active, Not Active, Not Active",
        "garbage",
    ]
    c = s.stop()
    assert c["sm_mhz"] == 1320.0 and c["sm_max_mhz"] == 1965.0 and c["reasons"] == ["sw_power_cap"]
    assert c["samples"] == 3 and c["power_w_max"] == pytest.approx(998.7)
    s2 = bench.ClockSampler(0)
    assert s2.stop()["reasons"] == ["nvidia-smi unavailable"]      # never started: no nvidia-smi, still a dict
