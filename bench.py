#!/usr/bin/env python
"""Benchmark of the policy-update hot path (BASELINE.json metric: policy-update tokens/sec/GPU for
loss head + advantage + loss, fwd/bwd, and % of roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|tiny]

One "step" = one pass of the hot path over one synthetic rollout batch:
    GRPO advantages -> fused LM head forward (log-probs + entropy) -> policy loss fwd/bwd (+KL term)
    -> fused LM head backward (d_hidden, d_weight).
Default workload = BASELINE config 2 (Qwen2.5-7B head: hidden 3584, vocab 152064; 256 seqs x 4096 tokens, group 8,
bf16) per GPU.  N > 1 (launched by torchrun, one rank per GPU) is weak scaling: every rank owns its own
prompt-group shard of that size; per step the packed stats vector and the fp32 lm_head weight gradient are
all-reduced over NCCL.  Rank 0 prints ONE JSON line.

Besides the contract keys the line carries
  `roofline`          the SLOWEST of the step's three GEMM kernels as the step runs them - its algorithmic flops per
                      launch over its average launch duration, measured LIVE with CUDA events around every launch
                      inside the timed steps (skyb200_policy_step_timing) - and `roofline_kernels` = all three;
                      `traffic` = dram bytes per launch from the committed ncu capture (profiles/r02_traffic.json),
                      null when the kernel sources changed after the capture;
  `step_roofline`     6*H*V*N algorithmic flops over the step time (the headline fraction);
  `hbm_kernels`       the memory-bound kernels of the path alone, L2 flushed, vs the measured HBM copy peak;
  `cpu_baseline`      the UNMODIFIED reference (baseline/_ref, kind "reference") on this box's host cores;
  `ref_gpu_path`      the UNMODIFIED reference's own GPU path (cuBLAS lm_head + flash-attn Triton CE + chunked entropy +
                      registry loss) on this GPU, and `vs_ref_gpu_path` = value / that;
  `other_configs`     BASELINE configs 1, 3, 4 (dual-clip, GSPO), 5 (per-rank batches, and one global ragged batch dealt
                      to the ranks by the prompt-group sharder) and the fixed-size config-3 split (strong scaling);
  `e2e`, `gpu_launches`, `clocks`; for N > 1 also `dp_check`, `allreduce_exposed_ms`, `per_rank_step_ms`.

`--impl reference` times the reference's own CPU implementation of the path (baseline/ref_path.py drives the
unmodified skyrl_train functions; oracle/ restatement only if no copy of the reference is present) on a bounded
token sample with all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

WORKLOADS = {
    # name: (B, A, H, V, group)
    "c2": dict(B=256, A=4096, H=3584, V=152064, G=8, model="Qwen2.5-7B head"),
    "c1": dict(B=64, A=1024, H=896, V=151936, G=8, model="Qwen2.5-0.5B head"),
    "c2s": dict(B=16, A=4096, H=3584, V=152064, G=8, model="Qwen2.5-7B head (16-sequence slice = 3 full chunks + a tail, profiling)"),
    "tiny": dict(B=16, A=256, H=256, V=4096, G=8, model="tiny"),
}
METRIC = "policy-update tokens/sec/GPU (loss head+adv+loss fwd/bwd)"
TEMPERATURE = 1.0


def measured_peaks():
    """Roofline denominators: the driver-written measurement of this pool's B200s, else (per key) the fallback
    B200_PROFILING.md states - the `source` string says which."""
    fallback = dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0)
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        d = json.load(open(p))
    except (OSError, ValueError):
        return dict(fallback, source="fallback (B200_PROFILING.md)")
    keys = dict(tflops="bf16_tflops_sustained", tflops_burst="bf16_tflops", hbm="hbm_gbs")
    out, missing = {}, []
    for k, name in keys.items():
        try:
            out[k] = float(d[name])
        except (KeyError, TypeError, ValueError):
            out[k] = fallback[k]
            missing.append(name)
    out["source"] = "measured (MEASURED_PEAKS.json, sustained bf16)" + (
        "; fallback (B200_PROFILING.md) for " + ", ".join(missing) if missing else "")
    return out


def config_dict(w, world):
    """The `config` object both arms print (same workload, same string)."""
    N = w["B"] * w["A"]
    return {
        "workload": f"{w['model']}: GRPO + regular loss (token_mean, k3 KL loss), {w['B']} seqs x {w['A']} tokens per GPU, "
                    f"H={w['H']}, V={w['V']}, group {w['G']}",
        "per_gpu_tokens": N,
        "l2": "inputs exceed L2 (hidden is %.1f GB per GPU)" % (N * w["H"] * 2 / 1e9),
        "parallelism": f"dp{world} by whole prompt groups",
    }


def make_cfg():
    from types import SimpleNamespace

    return SimpleNamespace(
        policy_loss_type="regular", loss_reduction="token_mean", eps_clip_low=0.2, eps_clip_high=0.2,
        clip_ratio_c=3.0, max_seq_len=None, use_kl_loss=True, kl_estimator_type="k3", kl_loss_coef=1e-3,
        use_entropy_loss=False, entropy_loss_coef=0.01,
        off_policy_correction=SimpleNamespace(tis_ratio_type=None, sequence_mask_metric=None),
    )


def host_batch(w, seed, pinned, tokens=None):
    """Synthetic rollout batch on the host.  Dense: every response position is valid."""
    B, A, H, V, G = w["B"], w["A"], w["H"], w["V"], w["G"]
    g = torch.Generator().manual_seed(seed)
    pin = dict(pin_memory=True) if pinned else {}
    hidden = torch.empty((B, A, H), dtype=torch.bfloat16, **pin)
    # fill in slabs to bound the fp32 temporary
    rows = max(1, (1 << 24) // (A * H))
    for b0 in range(0, B, rows):
        hidden[b0:b0 + rows] = torch.randn((min(rows, B - b0), A, H), generator=g).to(torch.bfloat16)
    labels = torch.empty((B, A), dtype=torch.int64, **pin).copy_(torch.randint(0, V, (B, A), generator=g))
    olp = torch.empty((B, A), dtype=torch.float32, **pin).copy_(-12.0 + 0.05 * torch.randn(B, A, generator=g))
    base = torch.empty((B, A), dtype=torch.float32, **pin).copy_(-12.0 + 0.1 * torch.randn(B, A, generator=g))
    mask = torch.ones((B, A), dtype=torch.float32, **pin)
    rewards = torch.zeros((B, A), dtype=torch.float32, **pin)
    rewards[:, -1] = (torch.rand(B, generator=g) < 0.3).float()
    uids = np.array([f"prompt-{seed}-{i // G}" for i in range(B)])
    return dict(hidden=hidden, labels=labels, olp=olp, base=base, mask=mask, rewards=rewards, uids=uids)


def make_weight(w, device):
    g = torch.Generator(device=device).manual_seed(99)
    V, H = w["V"], w["H"]
    out = torch.empty((V, H), dtype=torch.bfloat16, device=device)
    step = 16384
    for v0 in range(0, V, step):
        n = min(step, V - v0)
        out[v0:v0 + n] = (torch.randn((n, H), device=device, generator=g) / H ** 0.5).to(torch.bfloat16)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i",
                 str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_step(w, sample_tokens, threads, seed=1234):
    """The reference's torch CPU path on a bounded sample: (seconds, tokens, kind).  kind "reference" = the unmodified
    skyrl_train functions (baseline/ref_path.py), "port" = the oracle restatement when no copy is present."""
    from baseline import ref_path

    return ref_path.cpu_step(w, make_cfg(), sample_tokens, threads, TEMPERATURE, seed)


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = 1024 if w["V"] > 100000 else 4096
    times = []
    for i in range(args.warmup + args.steps):
        dt, n_tok, kind = cpu_reference_step(w, sample, threads, seed=1234 + i)
        if i >= args.warmup:
            times.append(dt)
    t = float(np.mean(times))
    value = n_tok / t
    what = ("unmodified skyrl_train functions (baseline/_ref) driven as worker.py:787-877" if kind == "reference"
            else "oracle/policy_oracle.py restatement (no copy of the reference present)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(w, args.gpus),
        "cpu_baseline": {"value": value, "unit": "tokens/s", "cores": threads, "kind": kind,
                         "sample": f"{n_tok} tokens/step of the same workload (head+loss fwd/bwd fp32, GRPO advantages "
                                   f"pro-rated); {what}"},
        "e2e": {"value": value, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------- GPU arm
def hbm_kernels(dev, peaks, B=512, flush_l2=True):
    """The memory-bound kernels of the path (K2 advantages, KL-in-reward, K3 loss) through the C ABI on the
    config-3 batch shape (512 x 8192 = 4.19 M positions, 50-117 MB per call): algorithmic bytes / CUDA-event time,
    L2 flushed before every timed call, median of 15.  Reported against the measured HBM copy bandwidth."""
    import ctypes as C

    from skyrl_b200 import _lib

    lib = _lib.load()
    A, G = 8192, 8
    n = B * A
    g = torch.Generator(device=dev).manual_seed(7)
    f = lambda: torch.randn((B, A), generator=g, device=dev)
    rewards, values, lp, adv = f(), f(), -f().abs(), f()
    olp, base = lp + 0.05 * f(), lp + 0.1 * f()
    lens = torch.randint(A // 2, A + 1, (B,), generator=g, device=dev)
    mask = (torch.arange(A, device=dev)[None] < lens[:, None]).float().contiguous()
    out1, out2 = torch.empty_like(lp), torch.empty_like(lp)
    stats = torch.zeros(64, dtype=torch.float32, device=dev)
    gid = (torch.arange(B, device=dev) // G).int()
    gstart = (torch.arange(B // G + 1, device=dev) * G).int()
    grows = torch.arange(B, device=dev).int()
    scores = torch.empty(4 * B, dtype=torch.float32, device=dev)
    gae_ws = torch.empty(lib.skyb200_gae_workspace(B), dtype=torch.uint8, device=dev)
    loss_ws = torch.empty(lib.skyb200_policy_loss_workspace(B, A), dtype=torch.uint8, device=dev)
    params = _lib.LossParams()
    params.loss_type, params.reduction, params.kl_type = 0, 0, 3
    params.eps_clip_low = params.eps_clip_high = 0.2
    params.clip_ratio_c, params.kl_loss_coef = 3.0, 1e-3
    st = _lib.stream_ptr()
    P = _lib.ptr
    cases = {
        # name: (call, algorithmic bytes per position)
        # uniform groups of 8 (what compute_grpo_outcome_advantage launches for this batch): one cluster launch
        "grpo_advantage": (lambda: lib.skyb200_group_advantage_uniform(P(rewards), P(mask), _lib.MASK_F32, P(grows),
                                                                       P(out1), P(scores), B, A, G, 0, 1e-6, st), 12),
        # the general form (ragged / shuffled groups): row sums, then group statistics + broadcast
        "grpo_advantage_general_form": (lambda: lib.skyb200_group_advantage(P(rewards), P(mask), _lib.MASK_F32, P(gid),
                                                                            P(gstart), P(grows), P(out1), P(scores),
                                                                            B, A, B // G, 0, 1e-6, st), 12),
        "gae_advantage+whiten": (lambda: lib.skyb200_gae_advantage(P(rewards), P(values), P(mask), _lib.MASK_F32,
                                                                   P(out1), P(out2), P(stats), P(gae_ws),
                                                                   gae_ws.numel(), B, A, 1.0, 0.95, st), 28),
        "reward_kl_penalty": (lambda: lib.skyb200_reward_kl_penalty(P(lp), P(base), P(mask), P(rewards), P(out1),
                                                                    P(stats), P(scores), B, A, 3, 1e-3, st), 20),
        "policy_loss_fwd_bwd(+kl)": (lambda: lib.skyb200_policy_loss(P(lp), P(olp), P(adv), P(mask), P(base), None,
                                                                     P(out1), None, P(stats), P(loss_ws),
                                                                     loss_ws.numel(), B, A, C.byref(params), None,
                                                                     None, st), 24),
    }
    # L2 flush = streaming READ of 3x the 126 MB L2 (a write flush would leave ~100 MB of dirty lines whose write-back
    # lands inside the timed region of a 50-100 MB kernel)
    flush = torch.zeros((96 << 20) if flush_l2 else 1, dtype=torch.int32, device=dev)
    res = {}
    def queued_times(fn, reps=15, skip=3):
        """Device time of `fn` between two CUDA events, `reps` times, WITHOUT a host sync in between: the stream is
        kept full (the L2 flush in front of every call takes ~70 us of GPU time, the host needs ~25 us per round), so
        the events bracket what the call costs inside an asynchronously enqueued training step - kernels plus the gaps
        between them - and not the host's launch latency after an idle GPU."""
        evs = []
        for _ in range(reps + skip):
            flush.sum()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            fn()
            e.record()
            evs.append((s, e))
        torch.cuda.synchronize()
        t = sorted(s.elapsed_time(e) for s, e in evs[skip:])
        return t[len(t) // 2]

    res["_timer_floor_us"] = round(queued_times(lambda: None) * 1e3, 2)
    for name, (call, bpt) in cases.items():
        for _ in range(3):
            _lib.check(call(), name)
        ms = queued_times(lambda: _lib.check(call(), name))
        gbs = n * bpt / (ms * 1e-3) / 1e9
        res[name] = {"us": round(ms * 1e3, 1), "bytes_per_position": bpt, "achieved_gbs": round(gbs, 1),
                     "frac_of_hbm_peak": round(gbs / peaks["hbm"], 3)}
    # calibration: what a one-launch library elementwise kernel reaches at this problem size with the same timing
    # method (the kernels above move 50-117 MB: launch ramp and tail are a visible share of 20-40 us)
    try:
        a, b, c = rewards, values, out1
        ms = queued_times(lambda: torch.add(a, b, out=c))
        gbs = n * 12 / (ms * 1e-3) / 1e9
        res["_calibration_torch_add"] = {"us": round(ms * 1e3, 1), "bytes_per_position": 12,
                                         "achieved_gbs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peaks["hbm"], 3)}
    except Exception as exc:  # noqa: BLE001  (context only: never fail the bench over it)
        res["_calibration_torch_add"] = {"error": str(exc)[:200]}
    res["_note"] = ("%d x 8192 positions%s, fp32 mask, %s, median of 15, CUDA events in a stream that is kept full (no host sync between "
                    "rounds; `_timer_floor_us` = an empty region measured the same way); hbm peak %.1f GB/s (%s)"
                    % (B, " (config-3 shape)" if B == 512 else " (%dx the config-3 batch)" % (B // 512),
                       "L2 flushed (384 MB streamed through it) before each timed call" if flush_l2 else
                       "no flush: every array (%d MB) exceeds what L2 keeps of it" % (n * 4 >> 20),
                       peaks["hbm"], peaks["source"].split(" (")[0]))
    return res


KERNEL_SOURCES = ["gemm_sm100.cuh", "head.cu", "ptx.cuh", "tile_order.cuh"]


def kernel_source_hash():
    """sha256 over the sources of the GEMM kernels: ties an ncu capture under profiles/ to the code it measured."""
    import hashlib

    h = hashlib.sha256()
    for name in KERNEL_SOURCES:
        with open(os.path.join(ROOT, "skyrl_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def measured_traffic(H, V):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the three GEMMs, from the committed `ncu --set full`
    capture of THIS build (profiles/r02_traffic.json, written by scripts/ncu_traffic.py from the capture's CSV).
    Returns ({kind: bytes}, note).  A capture of other sources or another head shape is refused loudly (traffic null)."""
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    try:
        d = json.load(open(path))
    except (OSError, ValueError):
        return {}, "no ncu capture committed for this build (profiles/r02_traffic.json missing)"
    if d.get("source_hash") != kernel_source_hash():
        print(f"bench.py: profiles/r02_traffic.json was captured from other kernel sources "
              f"({d.get('source_hash')} != {kernel_source_hash()}): roofline.traffic = null", file=sys.stderr)
        return {}, "STALE: the kernel sources changed after the ncu capture in profiles/r02_traffic.json"
    if (d.get("H"), d.get("V")) != (H, V):
        return {}, f"capture is for H={d.get('H')}, V={d.get('V')}"
    return {k: v["dram_read_bytes"] + v["dram_write_bytes"] for k, v in d["kernels"].items()}, d.get("capture", "")


def run_ours(args, w):
    import torch.distributed as dist

    import skyrl_b200
    from skyrl_b200 import dist as sdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200: the policy-update kernels have no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = measured_peaks()
    B, A, H, V = w["B"], w["A"], w["H"], w["V"]
    N = B * A
    cfg = make_cfg()

    hb = host_batch(w, 1234 + rank, pinned=True)
    weight = make_weight(w, dev)
    d_weight = torch.zeros((V, H), dtype=torch.float32, device=dev)
    db = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in hb.items()}
    # old / reference log-probs sit next to the current policy's (near on-policy rollouts):
    # old = lp + N(0, 0.05^2), base = lp + N(0, 0.1^2), 0.1 % outliers of +-25 to hit the +-20 ratio clamps.
    # (untimed set-up: one no-grad head forward gives lp)
    with torch.no_grad():
        lp0 = skyrl_b200.fused_lm_head_logprobs(db["hidden"].view(N, H), weight, db["labels"].view(N)).view(B, A)
        gd = torch.Generator(device=dev).manual_seed(4321 + rank)
        olp_d = lp0 + 0.05 * torch.randn(B, A, device=dev, generator=gd)
        base_d = lp0 + 0.1 * torch.randn(B, A, device=dev, generator=gd)
        outl = torch.rand(B, A, device=dev, generator=gd) < 1e-3
        olp_d = torch.where(outl, lp0 + 25.0 * torch.sign(torch.randn(B, A, device=dev, generator=gd)), olp_d)
        db["olp"], db["base"] = olp_d.contiguous(), base_d.contiguous()
        hb["olp"].copy_(db["olp"])
        hb["base"].copy_(db["base"])
        del lp0, olp_d, base_d, outl
    torch.cuda.synchronize()

    ev = lambda: torch.cuda.Event(enable_timing=True)
    phase_ms = {"_ev": []}

    # N > 1: how the lm_head weight gradient is reduced over the data-parallel ranks.
    #   "scatter"   = fused into the d_weight GEMM: every epilogue piece is averaged straight into its owner's
    #                 [V / world, H] shard over NVLink peer memory (FSDP's reduce-scatter semantics), no collective
    #                 after the step; the packed-stats all-reduce doubles as the "all ranks done" barrier;
    #   "allreduce" = local fp32 [V, H] gradient, then NCCL all-reduce (mean) after the step.
    scatter = None
    if world > 1 and args.dw_reduce == "scatter":
        scatter = sdist.DWeightScatter(V, H, remote=os.environ.get("SKYB200_DWS_REMOTE", "tma"))

    def device_step(record):
        """Inputs already resident in HBM."""
        e0, e1, e2, e3 = ev(), ev(), ev(), ev()
        if scatter is not None:
            scatter.begin()
        e0.record()
        adv, _ = skyrl_b200.compute_grpo_outcome_advantage(db["rewards"], db["mask"], db["uids"])
        e1.record()
        if scatter is not None:
            out = skyrl_b200.policy_update_step(db["hidden"], weight, db["labels"], db["olp"], adv, cfg,
                                                loss_mask=db["mask"], base_log_probs=db["base"],
                                                temperature=TEMPERATURE, dw_scatter=scatter, sync=False, timing=record)
        else:
            out = skyrl_b200.policy_update_step(db["hidden"], weight, db["labels"], db["olp"], adv, cfg,
                                                loss_mask=db["mask"], base_log_probs=db["base"],
                                                temperature=TEMPERATURE, d_weight=d_weight, accumulate_d_weight=False,
                                                sync=False, timing=record)
        e2.record()
        if world > 1:
            sdist.all_reduce_packed_stats(out.stats)  # (scatter mode: also the barrier that completes the shards)
            if scatter is None:
                sdist.all_reduce_d_weight(out.d_weight)
        e3.record()
        if record:
            phase_ms["_ev"].append((e0, e1, e2, e3))
        return out

    def e2e_step():
        """Host buffers in, python-float status out: what a user of the plugin API pays per step."""
        # the small per-token arrays go up front; the 7.5 GB of hidden states stay in PINNED host memory (required for
        # the asynchronous copy) and are streamed in chunk by chunk inside the step (skyb200_policy_step_from_host),
        # overlapping the GEMMs
        d = {k: (v.to(dev, non_blocking=True) if (torch.is_tensor(v) and k != "hidden") else v) for k, v in hb.items()}
        if scatter is not None:
            scatter.begin()
        adv, _ = skyrl_b200.compute_grpo_outcome_advantage(d["rewards"], d["mask"], d["uids"])
        kw = dict(dw_scatter=scatter) if scatter is not None else dict(d_weight=d_weight, accumulate_d_weight=False)
        out = skyrl_b200.policy_update_step(d["hidden"], weight, d["labels"], d["olp"], adv, cfg,
                                            loss_mask=d["mask"], base_log_probs=d["base"], temperature=TEMPERATURE,
                                            sync=(world == 1), **kw)
        if world > 1:
            st = sdist.all_reduce_packed_stats(out.stats)
            if scatter is None:
                sdist.all_reduce_d_weight(out.d_weight)
            return st.tolist()
        return out.status

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, record=False):
        for _ in range(warmup):
            fn(False) if record else fn()
        barrier()
        phase_ms["_ev"] = []
        s, e = ev(), ev()
        s.record()
        for _ in range(steps):
            fn(True) if record else fn()
        e.record()
        barrier()
        ms = s.elapsed_time(e)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_step = timed(device_step, args.steps, args.warmup, record=True)
    clocks = sampler.stop() if rank == 0 else None
    evs = phase_ms["_ev"]
    adv_ms = float(np.mean([a.elapsed_time(b) for a, b, _, _ in evs])) if evs else None
    upd_ms = float(np.mean([b.elapsed_time(c) for _, b, c, _ in evs])) if evs else None
    comm_ms = float(np.mean([c.elapsed_time(d) for _, _, c, d in evs])) if evs else None
    gemm_times = skyrl_b200.step_gemm_times(dev)  # live per-kernel durations of the timed steps
    e2e_steps = max(1, min(args.steps, 3))
    ms_e2e = timed(e2e_step, e2e_steps, 1)

    # the token chunk and the tile widths the library runs this shape with (host-only planning calls)
    import ctypes

    from skyrl_b200 import _lib

    lib = _lib.load()
    plan = (ctypes.c_int64 * 3)()
    _lib.check(lib.skyb200_policy_step_plan(B, A, 0, H, V, _lib.LOSS_TYPES["regular"], 1, plan), "skyb200_policy_step_plan")
    chunk, n_chunks = int(plan[0]), int(plan[2])
    tiles = (ctypes.c_int32 * 4)()
    _lib.check(lib.skyb200_head_tile_plan(min(chunk, N), H, V, tiles), "skyb200_head_tile_plan")
    # GRPO (1: the cluster kernel, uniform prompt groups) + loss pre-pass + finalize (2) + per chunk: fwd GEMM, merge,
    # loss, d_logits transform, dH GEMM, dW GEMM, live-row lists and dead-tile zero fill of the zero-row skipping
    # (profiles/r02t_launches_c2_step.csv: 450 launches between the pre-pass and the finalize of one step)
    launches_per_step = 1 + 2 + 8 * n_chunks

    # ---- multi-GPU evidence: N-GPU numerics + what the collectives cost with the rank skew taken out
    dp = None
    if world > 1:
        out = device_step(False)  # d_weight (or the shards) now hold the rank-averaged gradient of one step
        fused_err = None
        if scatter is not None:
            # fused reduce-scatter vs NCCL's mean reduce-scatter of the local gradients (computed below)
            fused_shard = scatter.finish().clone()
            adv, _ = skyrl_b200.compute_grpo_outcome_advantage(db["rewards"], db["mask"], db["uids"])
            skyrl_b200.policy_update_step(db["hidden"], weight, db["labels"], db["olp"], adv, cfg, loss_mask=db["mask"],
                                          base_log_probs=db["base"], temperature=TEMPERATURE, d_weight=d_weight,
                                          accumulate_d_weight=False, sync=False)
            want_shard, _ = sdist.reduce_scatter_d_weight(d_weight)
            fused_err = float((fused_shard.double() - want_shard.double()).norm() / want_shard.double().norm())
            del fused_shard, want_shard
            sdist.all_reduce_d_weight(d_weight)  # continue the common check on the all-reduced local gradient
        red_sum = d_weight.double().sum()
        red_abs = d_weight.double().abs().sum()
        red_loss = sdist.all_reduce_packed_stats(out.stats)[skyrl_b200._lib.STAT_FINAL_LOSS].double()
        adv, _ = skyrl_b200.compute_grpo_outcome_advantage(db["rewards"], db["mask"], db["uids"])
        loc = skyrl_b200.policy_update_step(db["hidden"], weight, db["labels"], db["olp"], adv, cfg,
                                            loss_mask=db["mask"], base_log_probs=db["base"], temperature=TEMPERATURE,
                                            d_weight=d_weight, accumulate_d_weight=False, sync=False)
        mine = torch.stack([d_weight.double().sum(), loc.stats[skyrl_b200._lib.STAT_FINAL_LOSS].double()])
        allv = torch.empty(world, 2, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allv, mine)
        want_sum, want_loss = allv[:, 0].mean(), allv[:, 1].mean()
        torch.cuda.synchronize()
        times = []
        for _ in range(3):  # the all-reduce alone, all ranks released together: NCCL time without the rank skew
            barrier()
            s, e = ev(), ev()
            s.record()
            sdist.all_reduce_d_weight(d_weight)
            e.record()
            torch.cuda.synchronize()
            times.append(s.elapsed_time(e))
        shard = torch.empty((V // world, H), dtype=torch.float32, device=dev) if V % world == 0 else None
        rs_times = []
        for _ in range(3 if shard is not None else 0):
            barrier()
            s, e = ev(), ev()
            s.record()
            sdist.reduce_scatter_d_weight(d_weight, out=shard)
            e.record()
            torch.cuda.synchronize()
            rs_times.append(s.elapsed_time(e))
        t = torch.tensor([min(times), min(rs_times) if rs_times else 0.0, upd_ms + adv_ms, comm_ms], device=dev,
                         dtype=torch.float64)
        tall = torch.empty(world, 4, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(tall, t)
        tall = tall.cpu()
        dp = {
            # checksum of the reduced gradient == mean of the per-rank checksums; reduced loss == mean of per-rank losses
            "d_weight_checksum_rel_err": float((red_sum - want_sum).abs() / red_abs.clamp(min=1e-30)),
            "final_loss_rel_err": float((red_loss - want_loss).abs() / want_loss.abs().clamp(min=1e-30)),
            "semantics": "mean over ranks of per-rank losses / gradients (worker.py:909-914, SURVEY trap 11)",
            "allreduce_ms": float(tall[:, 0].max()), "reduce_scatter_ms": float(tall[:, 1].max()) or None,
            "per_rank_compute_ms": [round(float(x), 1) for x in tall[:, 2]],
            "per_rank_comm_wait_ms_in_step": [round(float(x), 1) for x in tall[:, 3]],
        }
        dp["d_weight_reduction"] = ("fused into the d_weight GEMM epilogue: reduce-scatter over NVLink peer memory "
                                    "(DWeightScatter)" if scatter is not None else "NCCL all-reduce (mean) after the step")
        dp["fused_scatter_vs_nccl_reduce_scatter_rel_l2"] = fused_err
        dp["ok"] = bool(dp["d_weight_checksum_rel_err"] < 1e-5 and dp["final_loss_rel_err"] < 1e-5
                        and (fused_err is None or fused_err < 1e-5))

    # the library GEMM on the same shape and data, sustained (context for the power-capped regime)
    hid2d = db["hidden"].view(N, H)

    def cublas_same_shape():
        a = hid2d[:chunk]
        out = torch.empty((a.shape[0], V), dtype=torch.bfloat16, device=dev)
        for _ in range(3):
            torch.matmul(a, weight.t(), out=out)
        torch.cuda.synchronize()
        s, e = ev(), ev()
        reps = 0
        s.record()
        t_end = time.perf_counter() + 1.5
        while time.perf_counter() < t_end:
            for _ in range(10):
                torch.matmul(a, weight.t(), out=out)
            reps += 10
            torch.cuda.synchronize()
        e.record()
        torch.cuda.synchronize()
        return 2.0 * a.shape[0] * H * V * reps / (s.elapsed_time(e) * 1e-3) / 1e12

    cublas_tflops = cublas_same_shape() if rank == 0 else None
    hbm = hbm_kernels(dev, peaks) if rank == 0 else None
    # the same kernels on 4x the batch: launch ramp / tail and the 2.6 us timer floor no longer dominate
    hbm4 = hbm_kernels(dev, peaks, B=2048) if rank == 0 else None
    tok_s = world * N / (ms_step * 1e-3)
    tok_s_e2e = world * N / (ms_e2e * 1e-3)
    step_tflops_alg = 6.0 * N * H * V / (ms_step * 1e-3) / 1e12
    h2d = sum(v.numel() * v.element_size() for v in hb.values() if torch.is_tensor(v))

    # ---- roofline: the three GEMMs as the step runs them (live CUDA-event durations), the slowest one is `roofline`
    traffic, traffic_note = measured_traffic(H, V)
    kernel_names = {
        "forward": "gemm_sm100_kernel<2,256,6,K-major,K-major,EpiLseStoreRel> (logits tile -> LSE partials + bf16 tile store)",
        "d_hidden": f"gemm_sm100_kernel<2,{tiles[1]},{6 if tiles[1] == 256 else 4},K-major,MN-major,EpiStoreBF16> (d_hidden = dS * W)",
        "d_weight": f"gemm_sm100_kernel<2,{tiles[2]},{6 if tiles[2] == 256 else 4},MN-major,MN-major,EpiTmaF32> "
                    "(d_weight += dS^T * hidden, fp32 TMA reduce-add)",
    }
    roofline_kernels = {}
    for kind, t in gemm_times.items():
        if t["launches"] == 0:
            continue
        ms_launch = t["ms"] / t["launches"]
        flops_launch = 2.0 * N * H * V * args.steps / t["launches"]  # every launch covers one token chunk
        ach = flops_launch / (ms_launch * 1e-3) / 1e12
        roofline_kernels[kind] = {"kernel": kernel_names[kind], "bound": "tensor", "achieved": ach,
                                  "peak": peaks["tflops"], "unit": "TFLOP/s", "frac": ach / peaks["tflops"],
                                  "traffic": traffic.get(kind), "flops_per_launch": flops_launch,
                                  "ms_per_launch": ms_launch, "launches": t["launches"],
                                  "share_of_step": t["ms"] / (ms_step * args.steps)}
    worst = min(roofline_kernels, key=lambda k: roofline_kernels[k]["frac"]) if roofline_kernels else None

    # ---- other BASELINE configs (2 timed steps each) and the reference's own GPU path on this GPU
    other = None
    if args.workload == "c2" and not args.no_other:
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import bench_configs

        del db, hid2d, d_weight, weight  # free the config-2 batch (16 GB) before the larger shapes
        torch.cuda.empty_cache()
        other = {}
        for name in bench_configs.DEFAULT_SET:
            try:
                other[name] = bench_configs.run_config(name, dev, rank, world, steps=2, warmup=1, peaks=peaks)
            except Exception as exc:  # noqa: BLE001  (never lose the headline line over a side measurement)
                other[name] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
            torch.cuda.empty_cache()
    ref_gpu = None
    if rank == 0 and not args.no_other:
        from baseline import ref_path

        ref_gpu = ref_path.gpu_path(w, cfg, dev, micro_tokens=min(8192, N), temperature=TEMPERATURE)

    if rank == 0:
        threads = os.cpu_count() or 1
        sample = 4096 if V > 100000 else 16384  # ~10-20 s of host work
        dt, n_tok, kind = cpu_reference_step(w, sample, threads)
        line = {
            "metric": METRIC, "value": tok_s, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": config_dict(w, world),
            "chunk_plan": {"tokens_per_chunk": chunk, "chunks_per_step": n_chunks,
                           "tile_columns": {"forward": tiles[0], "d_hidden": tiles[1], "d_weight": tiles[2]},
                           "d_weight_n_group": tiles[3]},
            "tokens_per_s_per_gpu": tok_s / world,
            "phases_ms": {"advantages": adv_ms, "head_fwd+loss+head_bwd": upd_ms, "collectives(+rank skew)": comm_ms},
            "cublas_same_shape_tflops": cublas_tflops,
            "hbm_kernels": hbm,
            "hbm_kernels_4x_batch": hbm4,
            "step_roofline": {"bound": "tensor", "algorithmic_flops_per_token": 6.0 * H * V,
                              "achieved": step_tflops_alg, "peak": peaks["tflops"], "unit": "TFLOP/s",
                              "frac": step_tflops_alg / peaks["tflops"],
                              "note": "6*H*V*N algorithmic flops / step time (fused step: the backward reads the bf16 "
                                      "tile store, 6*H*V*N executed); peak = " + peaks["source"]},
            "roofline": dict(roofline_kernels[worst], which=worst,
                             note="slowest of the step's three GEMMs; duration = CUDA events around every launch "
                                  "inside the timed steps (live, power-capped regime); traffic: " + traffic_note
                                  + "; peak = " + peaks["source"]) if worst else None,
            "roofline_kernels": roofline_kernels,
            "cpu_baseline": {"value": n_tok / dt, "unit": "tokens/s", "cores": threads, "kind": kind,
                             "sample": f"{n_tok} tokens of the same workload, fp32 torch CPU path of the "
                                       + ("unmodified reference (baseline/_ref)" if kind == "reference"
                                          else "oracle restatement")},
            "ref_gpu_path": ref_gpu,
            "vs_ref_gpu_path": (tok_s / world / ref_gpu["tokens_per_s"]) if ref_gpu and "tokens_per_s" in ref_gpu else None,
            "other_configs": other,
            "e2e": {"value": tok_s_e2e, "unit": "tokens/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": 64, "ms_per_step": ms_e2e, "steps": e2e_steps,
                    "api": "compute_grpo_outcome_advantage + policy_update_step on PINNED host tensors (hidden streamed "
                           "in chunk by chunk inside the step; pageable memory would serialise the copy); d_hidden and "
                           "log-probs stay on the device for the trunk backward"},
            "gpu_launches": launches_per_step * args.steps,
            "dp_check": dp,
            "allreduce_exposed_ms": (0.0 if scatter is not None else dp["allreduce_ms"]) if dp else None,
            "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-other", action="store_true", help="skip other_configs / ref_gpu_path (A/B runs, profiling)")
    ap.add_argument("--dw-reduce", default="allreduce", choices=["allreduce", "scatter"],
                    help="N > 1: NCCL all-reduce of the local d_weight after the step, or the reduce-scatter fused into "
                         "the d_weight GEMM epilogue over NVLink peer memory")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
