"""Throughput of the fused policy step on the per-GPU shard of BASELINE configs 3, 4 and 5 (bench.py measures
config 2, the headline; these are the other named shapes).  One JSON line per config; under torchrun every rank owns
its own shard (weak scaling) and the packed stats + fp32 lm_head gradient are all-reduced per step, as in bench.py.

    python scripts/bench_configs.py [c1] [c3] [c3-strong] [c4-dual_clip] [c4-gspo] [c5] [c5-sharded] [steps=3] [warmup=2]

    c1  GRPO + regular loss + KL loss; 64 x 1024 tokens, Qwen2.5-0.5B head (H = 896, V = 151936) - the whole config
    c3-strong  config 3 at its FIXED size (512 x 8192 tokens) split over the ranks: strong scaling
    c3  PPO: KL-in-reward (k3) -> GAE (gamma 1, lambda 0.95) -> regular loss + KL loss; 64 x 8192 tokens, 7B head
    c4  dual-clip (c = 3, token_mean) or GSPO (sequence_mean); 8 x 32768-token trajectories (lengths 24K-32K),
        Qwen3-32B-shaped head (H = 5120, V = 151936)
    c5  RLOO (group 16) + entropy bonus; 64 ragged responses of 128-16K tokens, PACKED layout, 7B head
    c5-sharded  ONE global batch of 64 x world such responses (the same on every rank), whole prompt groups dealt to
        the ranks by dist.shard_prompt_groups (longest-processing-time first on the groups' valid tokens); reports the
        ranks' token counts next to what the reference's contiguous row split (dispatch.py:186-205) would give

`tokens_per_s` counts loss-masked (valid) tokens, `positions_per_s` every position the head processes; the roofline
fraction is 6*H*V flop per processed position against the measured sustained bf16 peak (zero-row skipping makes
masked positions cheaper in the backward GEMMs, so `frac` can exceed what the executed flops alone would give).
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import skyrl_b200  # noqa: E402
from bench import ClockSampler, make_cfg, make_weight, measured_peaks  # noqa: E402
from skyrl_b200 import dist as sdist  # noqa: E402

SHAPES = {
    "c1": dict(B=64, A=1024, H=896, V=151936, model="Qwen2.5-0.5B head"),
    "c3": dict(B=64, A=8192, H=3584, V=152064, model="Qwen2.5-7B head"),
    # the FIXED 512 x 8192 config-3 batch split over the ranks (strong scaling; B is per-rank = 512 / world)
    "c3-strong": dict(B=512, A=8192, H=3584, V=152064, model="Qwen2.5-7B head", strong=True),
    "c4-dual_clip": dict(B=8, A=32768, H=5120, V=151936, model="Qwen3-32B-shaped head"),
    "c4-gspo": dict(B=8, A=32768, H=5120, V=151936, model="Qwen3-32B-shaped head"),
    "c5": dict(B=64, A=16384, H=3584, V=152064, G=16, model="Qwen2.5-7B head"),
    "c5-sharded": dict(B=64, A=16384, H=3584, V=152064, G=16, model="Qwen2.5-7B head", sharded=True),
}


def _cfg(name):
    cfg = make_cfg()
    if name == "c4-dual_clip":
        cfg.policy_loss_type, cfg.eps_clip_high, cfg.clip_ratio_c, cfg.use_kl_loss = "dual_clip", 0.28, 3.0, False
    elif name == "c4-gspo":
        cfg.policy_loss_type, cfg.loss_reduction, cfg.use_kl_loss = "gspo", "sequence_mean", False
        cfg.eps_clip_low, cfg.eps_clip_high = 3e-4, 4e-4
    elif name.startswith("c5"):
        cfg.use_kl_loss, cfg.use_entropy_loss, cfg.entropy_loss_coef = False, True, 0.01
    return cfg


def _randn_bf16(n, h, gen, dev):
    out = torch.empty((n, h), dtype=torch.bfloat16, device=dev)
    for r0 in range(0, n, 65536):
        out[r0:r0 + 65536] = torch.randn((min(65536, n - r0), h), generator=gen, device=dev)
    return out


def build(name, dev, rank, world=1):
    """Returns (step_fn, n_valid_tokens, n_positions, description).  All inputs are resident in HBM."""
    s = SHAPES[name]
    B, A, H, V = s["B"], s["A"], s["H"], s["V"]
    if s.get("strong"):
        B = B // world
    name = "c3" if name == "c3-strong" else name
    shard_note = ""
    gen = torch.Generator(device=dev).manual_seed(1234 + 17 * rank)
    cfg = _cfg(name)
    weight = make_weight(s, dev)
    d_weight = torch.zeros((V, H), dtype=torch.float32, device=dev)
    ar = torch.arange(A, device=dev)[None]
    if name in ("c3", "c1"):
        lens = torch.full((B,), A, device=dev)
    elif name.startswith("c4"):
        lens = torch.randint(24 * 1024, A + 1, (B,), generator=gen, device=dev)
    elif s.get("sharded"):
        # the GLOBAL ragged batch, identical on every rank; this rank keeps the rows the sharder deals to it
        gg = torch.Generator().manual_seed(97531)
        lens_g = torch.randint(128, A + 1, (B * world,), generator=gg)
        uids_g = np.array([f"prompt-{i // s['G']}" for i in range(B * world)])
        shards = sdist.shard_prompt_groups(uids_g, lens_g.tolist(), world)
        per_rank = [int(lens_g[r].sum()) for r in shards]
        contiguous = [int(c.sum()) for c in lens_g.chunk(world)]
        imb = lambda v: max(v) / (sum(v) / len(v))
        shard_note = (f"; shard of a global batch of {B * world}: valid tokens per rank {per_rank} (max / mean "
                      f"{imb(per_rank):.3f}; the reference's contiguous split: {contiguous}, {imb(contiguous):.3f})")
        rows = shards[rank]
        lens, B, name = lens_g[rows].to(dev), len(rows), "c5"
        shard_uids = uids_g[rows]
    else:
        lens = torch.randint(128, A + 1, (B,), generator=gen, device=dev)
    mask = (ar < lens[:, None]).float()
    rewards = torch.zeros(B, A, device=dev)
    rewards[torch.arange(B, device=dev), lens - 1] = (torch.rand(B, generator=gen, device=dev) < 0.3).float()
    zeros = lambda *shape: torch.zeros(shape, device=dev)
    noise = lambda ref, sd: ref + sd * torch.randn(ref.shape, generator=gen, device=dev)

    if name == "c5":
        uids = shard_uids if s.get("sharded") else np.array([f"prompt-{rank}-{i // s['G']}" for i in range(B)])
        cu, (maskp,) = skyrl_b200.pack_by_lengths(lens, mask)
        N = int(cu[-1])
        hidden, labels = _randn_bf16(N, H, gen, dev), torch.randint(0, V, (N,), generator=gen, device=dev)
        lp0 = skyrl_b200.policy_update_step(hidden, weight, labels, zeros(N), zeros(N), cfg, loss_mask=maskp,
                                            need_d_hidden=False, need_d_weight=False, cu_seqlens=cu).log_probs
        olp = noise(lp0, 0.05)

        flat = torch.nonzero((ar < lens[:, None]).reshape(-1)).squeeze(-1)  # packed position -> padded position

        def step():
            adv, _ = skyrl_b200.compute_rloo_outcome_advantage(rewards, mask, uids)
            advp = adv.reshape(-1)[flat]
            return skyrl_b200.policy_update_step(hidden, weight, labels, olp, advp, cfg, loss_mask=maskp,
                                                 d_weight=d_weight, sync=False, cu_seqlens=cu)

        return (step, int(maskp.sum()), N,
                f"{B} ragged responses (128-16K tokens, {N} packed), RLOO group {s['G']}" + shard_note)

    N = B * A
    hidden, labels = _randn_bf16(N, H, gen, dev), torch.randint(0, V, (B, A), generator=gen, device=dev)
    lp0 = skyrl_b200.policy_update_step(hidden, weight, labels, zeros(B, A), zeros(B, A), cfg, loss_mask=mask,
                                        need_d_hidden=False, need_d_weight=False).log_probs
    if name == "c4-gspo":  # sequence log-ratios spread around the (tight) clip range
        olp = noise(lp0 - 3e-4 * torch.randn((B, 1), generator=gen, device=dev), 0.002)
    else:
        olp = noise(lp0, 0.05)
    base = noise(lp0, 0.1)
    if name == "c3":
        values = torch.randn((B, A), generator=gen, device=dev)

        def step():
            r_kl, _ = skyrl_b200.apply_reward_kl_penalty(rewards, lp0, base, mask, "k3", 1e-3)
            adv, _ = skyrl_b200.compute_gae_advantage_return(r_kl, values, mask, 1.0, 0.95)
            return skyrl_b200.policy_update_step(hidden, weight, labels, olp, adv, cfg, loss_mask=mask,
                                                 base_log_probs=base, d_weight=d_weight, sync=False)

        return step, int(mask.sum()), N, f"{B} x {A} tokens, KL-in-reward + GAE + regular loss + KL loss"
    uids = np.array([f"prompt-{rank}-{i // 8}" for i in range(B)])
    if name == "c1":

        def step():
            adv, _ = skyrl_b200.compute_grpo_outcome_advantage(rewards, mask, uids)
            return skyrl_b200.policy_update_step(hidden, weight, labels, olp, adv, cfg, loss_mask=mask,
                                                 base_log_probs=base, d_weight=d_weight, sync=False)

        return step, int(mask.sum()), N, f"{B} x {A} tokens, GRPO + regular loss + KL loss"

    def step():
        adv, _ = skyrl_b200.compute_grpo_outcome_advantage(rewards, mask, uids)
        return skyrl_b200.policy_update_step(hidden, weight, labels, olp, adv, cfg, loss_mask=mask,
                                             d_weight=d_weight, sync=False)

    return step, int(mask.sum()), N, f"{B} x {A}-token trajectories (lengths 24K-32K), GRPO + {cfg.policy_loss_type}"


DEFAULT_SET = ["c1", "c3", "c4-dual_clip", "c4-gspo", "c5", "c5-sharded", "c3-strong"]


def run_config(name, dev, rank, world, steps, warmup, peaks):
    """Time `steps` fused steps of config `name` (after `warmup`), CUDA events, max over ranks.  Returns the result
    dict on every rank (values are whole-job aggregates)."""
    import torch.distributed as dist

    step, n_valid, n_pos, desc = build(name, dev, rank, world)

    def full_step():
        out = step()
        if world > 1:
            sdist.all_reduce_packed_stats(out.stats)
            sdist.all_reduce_d_weight(out.d_weight)
        return out

    for _ in range(warmup):
        full_step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(dev.index)
    if rank == 0:
        sampler.start()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        out = full_step()
    e.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([s.elapsed_time(e) / steps, float(n_valid), float(n_pos)], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t[0] = tmax[0]
    ms, valid, pos = (float(x) for x in t.tolist())
    sh = SHAPES[name]
    tflops = 6.0 * sh["H"] * sh["V"] * pos / (ms * 1e-3) / 1e12 / world
    st = out.stats.tolist()
    res = {
        "config": name, "workload": f"{sh['model']} (H={sh['H']}, V={sh['V']}): {desc} per GPU",
        "n_gpus": world, "scaling": "strong" if sh.get("strong") else "weak", "steps": steps, "warmup": warmup,
        "ms_per_step": round(ms, 2),
        "tokens_per_s": round(valid / (ms * 1e-3)), "positions_per_s": round(pos / (ms * 1e-3)),
        "valid_fraction": round(valid / pos, 4),
        "step_roofline": {"bound": "tensor", "achieved": round(tflops, 1), "peak": peaks["tflops"],
                          "unit": "TFLOP/s per GPU", "frac": round(tflops / peaks["tflops"], 4),
                          "note": "6*H*V per processed position; peak = " + peaks["source"]},
        "final_loss": st[skyrl_b200._lib.STAT_FINAL_LOSS], "clip_ratio": st[skyrl_b200._lib.STAT_CLIP_RATIO],
        "clocks": clocks}
    del step, full_step, out
    torch.cuda.empty_cache()
    return res


def main():
    import torch.distributed as dist

    kv = dict(a.split("=", 1) for a in sys.argv[1:] if "=" in a)
    names = [a for a in sys.argv[1:] if "=" not in a] or [n for n in DEFAULT_SET]
    steps, warmup = int(kv.get("steps", 3)), int(kv.get("warmup", 2))
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = measured_peaks()
    for name in names:
        res = run_config(name, dev, rank, world, steps, warmup, peaks)
        if rank == 0:
            print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
