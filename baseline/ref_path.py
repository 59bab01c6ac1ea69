"""BASELINE ARMS ONLY - drives the UNMODIFIED reference (skyrl_train) over the policy-update path, on the host cores
(`bench.py --impl reference`, `cpu_baseline`) and on the GPU (`ref_gpu_path`: what SkyRL runs on this box today).

Nothing under ``skyrl_b200/`` imports this module, and nothing here touches ``libskyb200.so``.  The reference modules
are loaded by file path, unmodified, through ``oracle/ref_loader.py`` (stub ``ray`` / ``omegaconf`` /
``skyrl_train.config``) from ``/root/reference`` in the build container or from the copy ``__graft_entry__.build()``
vendors into ``baseline/_ref/`` (git-ignored; it travels to the GPU box with the repo snapshot).

The step mirrors ``PolicyWorkerBase._forward_backward_micro`` (workers/worker.py:787-877) after the trunk, calling the
reference's own functions:
    logits = lm_head(hidden); logits.div_(T)                                  model_wrapper.py:309-314
    logprobs_from_logits(logits, labels, inplace_backward=True)               utils/torch_utils.py:115-153
    chunked_entropy_from_logits(logits, requires_grad=use_entropy_loss)       utils/torch_utils.py:59-111
    PolicyLossRegistry.get(name)(lp, old_lp, adv, config, loss_mask=...)      utils/ppo_utils.py:548-586
    compute_approx_kl + masked_mean(dim=-1).mean()                            workers/worker.py:865-871
    loss = policy + kl * coef - entropy * coef; loss.backward()               workers/worker.py:876-877
and the advantage estimators through ``AdvantageEstimatorRegistry`` as the trainer calls them (trainer.py:789, 810).
"""
from __future__ import annotations

import os
import shutil
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

VENDOR_DIR = os.path.join(ROOT, "baseline", "_ref", "skyrl_train")
SOURCE_DIR = "/root/reference/skyrl-train/skyrl_train"
FILES = ["training_batch.py", "utils/torch_utils.py", "utils/off_policy_correction_utils.py", "utils/ppo_utils.py"]


def vendor() -> bool:
    """Copy the four hot-path files of the reference, unmodified, into baseline/_ref/ (SURVEY.md section 7, step 0).
    Returns False when /root/reference is absent (GPU box: the prebuilt copy is used)."""
    if not os.path.isfile(os.path.join(SOURCE_DIR, FILES[-1])):
        return False
    for rel in FILES:
        dst = os.path.join(VENDOR_DIR, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SOURCE_DIR, rel), dst)
    return True


def load():
    """(ppo_utils, torch_utils) of the unmodified reference, or None when neither copy exists."""
    from oracle import ref_loader

    if not ref_loader.available():
        return None
    return ref_loader.load()


def _head_loss_step(pu, tu, hidden, weight, labels, olp, base, adv, mask, cfg, temperature, on_gpu):
    """One micro-batch of worker.py:787-877 after the trunk.  hidden [rows*a, H], weight [V, H] (leaf tensors)."""
    import torch
    import torch.nn.functional as F

    rows, a_len = olp.shape
    logits = F.linear(hidden.view(rows, a_len, -1), weight)   # the HF lm_head: [batch, seqlen, vocab]
    logits.div_(temperature)                                  # model_wrapper.py:314
    labels = labels.view(rows, a_len)
    if on_gpu:
        lp = tu.logprobs_from_logits(logits, labels, inplace_backward=True)   # flash-attn Triton cross-entropy
    else:
        # flash_attn is importable in this image, so the reference's dispatcher would take the Triton branch, which
        # cannot run on CPU tensors; the reference's own CPU branch is logprobs_from_logits_v2 (torch_utils.py:158)
        lp = tu.logprobs_from_logits_v2(logits, labels)
    ent = tu.chunked_entropy_from_logits(logits, requires_grad=bool(cfg.use_entropy_loss))
    lp2 = lp.view(rows, a_len)
    loss_fn = pu.PolicyLossRegistry.get(cfg.policy_loss_type)
    policy_loss, metrics = loss_fn(lp2, olp, adv, config=cfg, loss_mask=mask, rollout_logprobs=None)
    with torch.set_grad_enabled(bool(cfg.use_entropy_loss)):
        entropy = tu.masked_mean(ent.view(rows, a_len), mask)
    ent_term = entropy * cfg.entropy_loss_coef if cfg.use_entropy_loss else torch.tensor(0.0)
    if cfg.use_kl_loss:
        kl = pu.compute_approx_kl(lp2, base, loss_mask=mask, kl_estimator_type=cfg.kl_estimator_type)
        kl_loss = tu.masked_mean(kl, mask, dim=-1).mean()
    else:
        kl_loss = torch.tensor(0.0)
    loss = policy_loss + kl_loss * cfg.kl_loss_coef - ent_term
    loss.backward()
    return loss, metrics


def cpu_step(w, cfg, sample_tokens, threads, temperature=1.0, seed=1234):
    """The reference's torch CPU path on a bounded sample of workload `w`: GRPO advantages over the full [B, A] batch
    (pro-rated to the sample) + head and loss fwd/bwd in fp32 on `sample_tokens` tokens.  Returns (seconds, tokens,
    kind) with kind = "reference" (unmodified reference) or "port" (oracle restatement, when no copy is present)."""
    import numpy as np
    import torch

    mods = load()
    torch.set_num_threads(threads)
    B, A, H, V, G = w["B"], w["A"], w["H"], w["V"], w["G"]
    g = torch.Generator().manual_seed(seed)
    rows = max(1, sample_tokens // A) if sample_tokens >= A else 1
    a_len = A if sample_tokens >= A else sample_tokens
    n_tok = rows * a_len
    hidden = torch.randn(n_tok, H, generator=g).to(torch.bfloat16).float()
    weight = (torch.randn(V, H, generator=g) / H ** 0.5).to(torch.bfloat16).float()
    labels = torch.randint(0, V, (n_tok,), generator=g)
    with torch.no_grad():  # untimed set-up, as in the GPU arm: old / base log-probs next to the current policy's
        lp0 = torch.cat([torch.log_softmax(hidden[s:s + 256] @ weight.t() / temperature, -1)
                         .gather(-1, labels[s:s + 256, None]).squeeze(-1) for s in range(0, n_tok, 256)])
    olp = lp0.view(rows, a_len) + 0.05 * torch.randn(rows, a_len, generator=g)
    base = lp0.view(rows, a_len) + 0.1 * torch.randn(rows, a_len, generator=g)
    mask = torch.ones(rows, a_len)
    rewards = torch.zeros(B, A)
    rewards[:, -1] = (torch.rand(B, generator=g) < 0.3).float()
    uids = np.array([f"p{i // G}" for i in range(B)])
    h = hidden.requires_grad_(True)
    wt = weight.requires_grad_(True)
    if mods is None:
        from oracle import policy_oracle as O

        t0 = time.perf_counter()
        adv_full, _ = O.grpo_advantage(rewards, torch.ones(B, A), uids)
        t_adv = time.perf_counter() - t0
        adv = adv_full[:rows, :a_len]
        t0 = time.perf_counter()
        lp, ent = O.head_logprobs_entropy(h, wt, labels, temperature, chunk=256)
        final, _ = O.loss_assembly(lp.view(rows, a_len), olp, adv, cfg, mask, base, ent.view(rows, a_len))
        final.backward()
        return time.perf_counter() - t0 + t_adv * (n_tok / (B * A)), n_tok, "port"
    pu, tu = mods
    t0 = time.perf_counter()
    adv_full, _ = pu.AdvantageEstimatorRegistry.get("grpo")(
        token_level_rewards=rewards, response_mask=torch.ones(B, A), index=uids, values=None, grpo_norm_by_std=True,
        gamma=1.0, lambd=1.0, config=cfg)
    t_adv = time.perf_counter() - t0
    adv = adv_full[:rows, :a_len].contiguous()
    t0 = time.perf_counter()
    _head_loss_step(pu, tu, h, wt, labels, olp, base, adv, mask, cfg, temperature, on_gpu=False)
    t_head = time.perf_counter() - t0
    # advantages are O(B*A) for the whole batch; pro-rate them to the sample so the figure is per-token consistent
    return t_head + t_adv * (n_tok / (B * A)), n_tok, "reference"


def gpu_path(w, cfg, dev, micro_tokens=8192, reps=5, temperature=1.0):
    """The reference's own GPU path for the same work on this GPU ("the real bar to beat"): bf16
    lm_head (cuBLAS) -> div_ -> flash-attn Triton cross-entropy (in-place backward) -> chunked entropy -> registry
    loss + KL term -> backward, under torch.autocast(bf16) like the worker, on a micro-batch of `micro_tokens`
    tokens ([tokens x vocab] logits must fit: 8192 x 152064 bf16 = 2.5 GB + the fp32 temporaries of the entropy).
    Returns a dict (tokens_per_s, ...) or {"unavailable": why}."""
    import torch

    mods = load()
    if mods is None:
        return {"unavailable": "no copy of the reference (baseline/_ref is built by __graft_entry__.build())"}
    pu, tu = mods
    try:
        import flash_attn  # noqa: F401  (the reference's logprobs_from_logits dispatches on it)
    except Exception as exc:  # noqa: BLE001
        return {"unavailable": f"flash_attn not importable: {exc}"[:200]}
    H, V, A = w["H"], w["V"], w["A"]
    g = torch.Generator(device=dev).manual_seed(0)
    rows = max(1, micro_tokens // A)
    a_len = micro_tokens // rows
    n_tok = rows * a_len
    hidden = torch.randn(n_tok, H, device=dev, generator=g).bfloat16().requires_grad_(True)
    weight = (torch.randn(V, H, device=dev, generator=g) / H ** 0.5).bfloat16().requires_grad_(True)
    labels = torch.randint(0, V, (n_tok,), device=dev, generator=g)
    olp = -12.0 + 0.05 * torch.randn(rows, a_len, device=dev, generator=g)
    base = -12.0 + 0.1 * torch.randn(rows, a_len, device=dev, generator=g)
    adv = torch.randn(rows, a_len, device=dev, generator=g)
    mask = torch.ones(rows, a_len, device=dev)

    def step():
        hidden.grad = None
        weight.grad = None
        with torch.autocast(dtype=torch.bfloat16, device_type="cuda"):   # worker.py:788
            return _head_loss_step(pu, tu, hidden, weight, labels, olp, base, adv, mask, cfg, temperature, on_gpu=True)

    try:
        for _ in range(2):
            step()
        torch.cuda.synchronize(dev)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            loss, _ = step()
        e.record()
        torch.cuda.synchronize(dev)
    except Exception as exc:  # noqa: BLE001
        return {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}
    ms = s.elapsed_time(e) / reps
    out = {"what": "UNMODIFIED skyrl_train functions on this GPU: bf16 F.linear + logits.div_ + logprobs_from_logits "
                   "(flash-attn Triton CE, inplace backward) + chunked_entropy_from_logits + PolicyLossRegistry['%s'] "
                   "+ compute_approx_kl + backward, under autocast(bf16)" % cfg.policy_loss_type,
           "micro_batch_tokens": n_tok, "ms_per_micro_batch": ms, "tokens_per_s": n_tok / (ms * 1e-3),
           "final_loss": float(loss), "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9}
    del hidden, weight
    torch.cuda.empty_cache()
    return out
