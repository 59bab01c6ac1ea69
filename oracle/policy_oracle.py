"""CPU (torch fp32) restatement of the reference's policy-update hot path.  TEST INFRASTRUCTURE — see
``oracle/__init__.py``.  Each function names the reference lines it restates (paths relative to
``skyrl-train/skyrl_train/``).  Written vectorised where the reference loops in Python; the sequential fp32
accumulation order of the reference is kept where it matters for parity (the GAE recurrence).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, Optional, Sequence, Tuple

import math

import numpy as np
import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------- small helpers
def masked_mean(x: torch.Tensor, mask: Optional[torch.Tensor], dim=None) -> torch.Tensor:
    """utils/torch_utils.py:180-184 — denominator clamped to >= 1."""
    if mask is None:
        return x.mean(dim) if dim is not None else x.mean()
    num = (x * mask).sum(dim) if dim is not None else (x * mask).sum()
    den = mask.sum(dim) if dim is not None else mask.sum()
    return num / den.clamp(min=1.0)


def masked_var(x: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """utils/ppo_utils.py:148-160 — Bessel-corrected; raises when the mask sums to 0 or 1."""
    mu = masked_mean(x, mask)
    var = masked_mean((x - mu) ** 2, mask)
    n = mask.sum()
    if n == 0:
        raise ValueError("At least one element in the mask has to be 1.")
    if n == 1:
        raise ValueError("The sum of the mask is one, which can cause a division by zero.")
    return var * (n / (n - 1))


def masked_whiten(x: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """utils/ppo_utils.py:163-172 (shift_mean=True)."""
    mu, var = masked_mean(x, mask), masked_var(x, mask)
    return (x - mu) * torch.rsqrt(var + 1e-8)


def approx_kl(lp: torch.Tensor, base: torch.Tensor, mask: Optional[torch.Tensor], kind: str = "k3") -> torch.Tensor:
    """utils/ppo_utils.py:88-124 (no gradient)."""
    with torch.no_grad():
        d = lp - base
        if kind == "k1":
            out = d
        elif kind == "abs":
            out = d.abs()
        elif kind == "k2":
            out = 0.5 * d * d
        elif kind == "k3":
            x = (-d).clamp(-20.0, 20.0)
            out = (x.exp() - x - 1.0).clamp(-10.0, 10.0)
        else:
            raise ValueError(f"Invalid KL estimator type: {kind}")
        return out * mask if mask is not None else out


# ---------------------------------------------------------------------------------------- advantages
def _group_ids(index: Sequence) -> Tuple[torch.Tensor, int]:
    table: Dict = {}
    ids = []
    for u in index:
        u = u.item() if isinstance(u, np.generic) else u
        ids.append(table.setdefault(u, len(table)))
    return torch.tensor(ids, dtype=torch.int64), len(table)


def grpo_advantage(rewards, response_mask, index, eps: float = 1e-6, norm_by_std: bool = True):
    """utils/ppo_utils.py:1132-1182 — unmasked row sums, per-uid mean / unbiased std, eps on std,
    singleton -> mean 0 std 1, result * mask; (adv, returns) are the same tensor."""
    scores = rewards.float().sum(-1)
    gid, ng = _group_ids(index)
    cnt = torch.zeros(ng).index_add_(0, gid, torch.ones_like(scores))
    mean = torch.zeros(ng).index_add_(0, gid, scores) / cnt
    dev2 = torch.zeros(ng).index_add_(0, gid, (scores - mean[gid]) ** 2)
    std = torch.sqrt(dev2 / (cnt - 1).clamp(min=1.0))
    single = cnt == 1
    mean = torch.where(single, torch.zeros_like(mean), mean)
    std = torch.where(single, torch.ones_like(std), std)
    out = (scores - mean[gid]) / (std[gid] + eps) if norm_by_std else scores - mean[gid]
    adv = out.unsqueeze(-1) * response_mask
    return adv, adv


def rloo_advantage(rewards, response_mask, index):
    """utils/ppo_utils.py:1048-1098 — (score - group mean) * n/(n-1); singleton -> 0."""
    scores = rewards.float().sum(-1)
    gid, ng = _group_ids(index)
    cnt = torch.zeros(ng).index_add_(0, gid, torch.ones_like(scores))
    mean = torch.zeros(ng).index_add_(0, gid, scores) / cnt
    n = cnt[gid]
    out = torch.where(n > 1, (scores - mean[gid]) * (n / (n - 1).clamp(min=1.0)), torch.zeros_like(scores))
    adv = out.unsqueeze(-1) * response_mask
    return adv, adv


def gae_advantage_return(rewards, values, response_mask, gamma: float, lambd: float):
    """utils/ppo_utils.py:1101-1129 — sequential fp32 recurrence over ALL positions (no mask reset),
    returns = raw advantages + values, advantages whitened over the whole batch, not re-masked."""
    r = rewards.float().numpy()
    v = values.float().numpy()
    B, A = r.shape
    adv = np.zeros((B, A), dtype=np.float32)
    last = np.zeros(B, dtype=np.float32)
    g = np.float32(gamma)
    gl = np.float32(gamma * lambd)
    for t in range(A - 1, -1, -1):
        nxt = v[:, t + 1] if t + 1 < A else np.zeros(B, dtype=np.float32)
        delta = r[:, t] + g * nxt - v[:, t]
        last = (delta + gl * last).astype(np.float32)
        adv[:, t] = last
    adv_t = torch.from_numpy(adv)
    returns = adv_t + values.float()
    return masked_whiten(adv_t, response_mask), returns


def reinforce_pp_advantage(rewards, response_mask, gamma: float):
    """utils/ppo_utils.py:1013-1045 — discounted returns, running return reset after EOS (running *= mask[t]),
    advantages = masked_whiten(returns, mask) * mask."""
    r = rewards.float().numpy()
    m = response_mask.float().numpy()
    B, A = r.shape
    ret = np.zeros((B, A), dtype=np.float32)
    run = np.zeros(B, dtype=np.float32)
    g = np.float32(gamma)
    for t in range(A - 1, -1, -1):
        run = (r[:, t] + g * run).astype(np.float32)
        ret[:, t] = run
        run = run * m[:, t]
    ret_t = torch.from_numpy(ret)
    mask = response_mask.float()
    return masked_whiten(ret_t, mask) * mask, ret_t


def normalize_advantages(adv, response_mask):
    """utils/ppo_utils.py:127-145 — mean over all positions, variance over the mask, rsqrt(clamp(var, 1e-8))."""
    mask = response_mask.float()
    mean = adv.mean()
    var = ((adv - mean).pow(2) * mask).sum() / mask.sum()
    return (adv - mean) * var.clamp(min=1e-8).rsqrt()


def critic_loss(values, old_values, returns, value_clip, mask=None):
    """utils/ppo_utils.py:175-193 — returns (0.5 * loss, clipfrac or None)."""
    if value_clip is not None:
        clipped = old_values + (values - old_values).clamp(-value_clip, value_clip)
        s1, s2 = (clipped - returns) ** 2, (values - returns) ** 2
        loss = torch.max(s1, s2)
        clipfrac = masked_mean((s1 > s2).float(), mask).mean().item()
    else:
        loss, clipfrac = (values - returns) ** 2, None
    return 0.5 * masked_mean(loss, mask, dim=-1).mean(), clipfrac


def reward_kl_penalty(rewards, lp, base, loss_mask, kind: str = "k3", coef: float = 1e-3):
    """trainer.py:981-1035 — rewards - kl*max(0, coef); avg_kl = mean_b masked_mean_t kl; avg_kl_max = mean_b max_t|kl|."""
    kl = approx_kl(lp, base, loss_mask, kind)
    kl_max = kl.abs().max(dim=-1)[0]
    kl_mean = masked_mean(kl, loss_mask, dim=-1)
    out = rewards - kl * max(0, coef)
    return out, {"avg_kl": kl_mean.mean().item(), "avg_kl_max": kl_max.mean().item()}


# ---------------------------------------------------------------------------------------- policy loss
def reduce_loss(loss, mask, reduction: str, max_seq_len=None):
    """utils/ppo_utils.py:984-1009."""
    if reduction == "token_mean":
        return masked_mean(loss, mask)
    if reduction == "sequence_mean":
        return masked_mean(loss, mask, dim=-1).mean()
    if reduction == "seq_mean_token_sum_norm":
        assert max_seq_len is not None
        tot = (loss * mask).sum(-1) if mask is not None else loss.sum(-1)
        return (tot / max_seq_len).mean()
    raise ValueError(f"Invalid loss reduction type: {reduction}")


def policy_loss(lp, olp, adv, cfg, mask=None):
    """"regular"/"dual_clip": utils/ppo_utils.py:548-586; "gspo": :661-725.  Returns (loss, clip_ratio)."""
    kind = cfg.policy_loss_type
    if kind == "gspo":
        seq = masked_mean(lp - olp, mask, dim=-1).unsqueeze(-1)
        log_w = (lp - lp.detach() + seq.detach()).clamp(max=10.0)
        ratio = log_w.exp()
    else:
        ratio = (lp - olp).float().clamp(-20.0, 20.0).exp()
    s1 = ratio * adv
    s2 = ratio.clamp(1 - cfg.eps_clip_low, 1 + cfg.eps_clip_high) * adv
    loss = -torch.min(s1, s2)
    clip_ratio = masked_mean((-s2 > -s1).float(), mask).mean().item()
    if kind == "dual_clip":
        capped = torch.min(-adv * cfg.clip_ratio_c, loss)
        loss = torch.where(adv < 0, capped, loss)
    return reduce_loss(loss, mask, cfg.loss_reduction, getattr(cfg, "max_seq_len", None)), clip_ratio


def off_policy_correction(olp, rlp, mask, opc):
    """utils/off_policy_correction_utils.py:173-262 restated: returns (tis or None, metrics, new_mask)."""
    tis_type, seq_metric = opc.tis_ratio_type, opc.sequence_mask_metric
    if tis_type is None and seq_metric is None:
        return None, {}, mask
    safe = lambda d: d.float().clamp(-20.0, 20.0).exp()
    d = olp - rlp
    ratio = safe(d)
    met = {"is_ratio_mean": masked_mean(ratio, mask).item(), "is_ratio_std": (ratio * mask).std().item(),
           "is_ratio_max": (ratio * mask).max().item(), "is_ratio_min": (ratio * mask).min().item()}
    hi, lo = opc.outlier_token_is_threshold_high, opc.outlier_token_is_threshold_low
    over = (ratio > hi) & (mask > 0) if hi is not None else torch.zeros_like(mask, dtype=torch.bool)
    under = (ratio < lo) & (mask > 0) if lo is not None else torch.zeros_like(mask, dtype=torch.bool)
    row_ok = ~(over.any(-1, keepdim=True) | under.any(-1, keepdim=True))
    B = float(mask.shape[0])
    met.update(outlier_seq_masked_ratio=((~row_ok).sum() / B).item(), outlier_seq_over_high_ratio=(over.any(-1).sum() / B).item(),
               outlier_seq_under_low_ratio=(under.any(-1).sum() / B).item())
    mask = mask * row_ok.float()
    tis = None
    if tis_type == "token":
        cap = opc.token_tis_ratio_clip_high
        met["tis_token_clip_high_ratio"] = (((ratio > cap) & (mask > 0)).sum() / (mask > 0).sum().clamp(min=1)).item()
        tis = ratio.clamp(max=cap)
    elif tis_type == "sequence":
        sr = safe((d * mask).sum(-1, keepdim=True))
        cap = opc.sequence_tis_ratio_clip_high
        met["tis_seq_clip_high_ratio"] = ((sr > cap).sum() / B).item()
        tis = sr.clamp(max=cap)
    elif tis_type is not None:
        raise ValueError(f"Unknown tis_ratio_type: {tis_type}")
    if seq_metric is not None:
        s = (d * mask).sum(-1, keepdim=True)
        if seq_metric == "geometric":
            val, lo_, hi_, nm = safe(s / mask.sum(-1, keepdim=True).clamp(min=1.0)), opc.geo_mask_low, opc.geo_mask_high, "geo"
        elif seq_metric == "product":
            val, lo_, hi_, nm = safe(s), opc.product_mask_low, opc.product_mask_high, "product"
        else:
            raise ValueError(f"Unknown sequence_mask_metric: {seq_metric}")
        so, su = val > hi_, val < lo_
        keep = ~so & ~su
        met[f"{nm}_sequence_mask_masked_ratio"] = ((~keep).sum() / B).item()
        met[f"{nm}_sequence_mask_over_high_ratio"] = (so.sum() / B).item()
        met[f"{nm}_sequence_mask_under_low_ratio"] = (su.sum() / B).item()
        mask = mask * keep.float()
    return tis, met, mask


def policy_loss_with_correction(lp, olp, adv, cfg, mask, rlp):
    """ppo_policy_loss with rollout_logprobs (utils/ppo_utils.py:548-586 incl. :579-583). Returns (loss, metrics)."""
    ratio = (lp - olp).float().clamp(-20.0, 20.0).exp()
    s1 = ratio * adv
    s2 = ratio.clamp(1 - cfg.eps_clip_low, 1 + cfg.eps_clip_high) * adv
    loss = -torch.min(s1, s2)
    metrics = {"clip_ratio": masked_mean((-s2 > -s1).float(), mask).mean().item()}
    if cfg.policy_loss_type == "dual_clip":
        loss = torch.where(adv < 0, torch.min(-adv * cfg.clip_ratio_c, loss), loss)
    tis, met, mask = off_policy_correction(olp, rlp, mask, cfg.off_policy_correction)
    if tis is not None:
        loss = loss * tis
    metrics.update(met)
    return reduce_loss(loss, mask, cfg.loss_reduction, getattr(cfg, "max_seq_len", None)), metrics


def extra_policy_loss(kind, lp, olp, adv, cfg, mask=None):
    """sapo (utils/ppo_utils.py:589-658), cispo (:728-761), cross_entropy (:892-930), importance_sampling (:933-981).
    Returns (loss, metrics)."""
    if kind == "sapo":
        ratio = (lp - olp).clamp(-20.0, 20.0).exp()
        tau = torch.where(adv > 0, torch.as_tensor(cfg.sapo.tau_pos, dtype=adv.dtype),
                          torch.as_tensor(cfg.sapo.tau_neg, dtype=adv.dtype))
        gate = torch.sigmoid(tau * (ratio - 1.0)) * (4.0 / tau)
        return reduce_loss(-gate * adv, mask, cfg.loss_reduction, getattr(cfg, "max_seq_len", None)), {"clip_ratio": 0.0}
    if kind == "cispo":
        ratio = (lp - olp).float().clamp(-20.0, 20.0).exp()
        lo, hi = 1 - cfg.cispo.cispo_eps_clip_low, 1 + cfg.cispo.cispo_eps_clip_high
        loss = -adv * ratio.clamp(lo, hi).detach() * lp
        clip = masked_mean(((ratio < lo) | (ratio > hi)).float(), mask).mean().item()
        return reduce_loss(loss, mask, cfg.loss_reduction, getattr(cfg, "max_seq_len", None)), {"clip_ratio": clip}
    if kind == "cross_entropy":
        return ((-lp * mask).sum() if mask is not None else (-lp).sum()), {"clip_ratio": 0.0}
    if kind == "importance_sampling":
        ratio = (lp - olp).exp()
        el = -(ratio * adv)
        if mask is not None:
            return (el * mask).sum(), {"importance_ratio": ((ratio * mask).sum() / mask.sum()).item()}
        return el.sum(), {"importance_ratio": ratio.mean().item()}
    raise ValueError(kind)


def cov_selection(kind, lp, olp, adv, cfg, mask, perm_fn=None):
    """The detached token selection of clip_cov (utils/ppo_utils.py:785-814) / kl_cov (:857-871) as a 0/1 tensor.

    clip_cov draws a random subset of the candidate tokens with ``torch.randperm`` (:806); ``perm_fn(n)`` replaces
    that draw (default: torch.randperm, so a seeded oracle replays the seeded reference)."""
    lp_d = lp.detach()
    sel = torch.zeros_like(adv)
    if kind == "clip_cov":
        c = cfg.clip_cov
        ratio = (lp_d - olp).exp()
        l1 = -adv * ratio
        l2 = -adv * ratio.clamp(1 - cfg.eps_clip_low, 1 + cfg.eps_clip_high)
        by_origin = (l2 > l1) & (mask > 0)
        cov = (adv - masked_mean(adv, mask)) * (lp_d - masked_mean(lp_d, mask))
        cov = cov.masked_fill(mask == 0, -math.inf).masked_fill(by_origin, -math.inf)
        clip_num = max(int(c.clip_ratio * mask.sum().item()), 1)
        cand = torch.nonzero((cov < c.clip_cov_ub) & (cov > c.clip_cov_lb) & (mask > 0))
        if len(cand) > 0:
            perm = (perm_fn or torch.randperm)(len(cand))
            cand = cand[perm[: min(clip_num, len(cand))]]
            sel[cand[:, 0], cand[:, 1]] = 1.0
        return sel
    if kind == "kl_cov":
        valid = mask > 0
        idx = torch.nonzero(valid.reshape(-1), as_tuple=True)[0]
        a, l = adv[valid].reshape(-1), lp_d[valid].reshape(-1)
        if len(a) > 0:
            cov = (a - a.mean()) * (l - l.mean())
            k = max(1, int(len(cov) * cfg.kl_cov.kl_cov_frac))
            top = idx[torch.topk(cov, min(k, len(cov)), largest=True).indices]
            sel.view(-1)[top] = 1.0
        return sel
    raise ValueError(kind)


def cov_policy_loss(kind, lp, olp, adv, cfg, mask, sel=None, perm_fn=None):
    """clip_cov (utils/ppo_utils.py:765-829) and kl_cov (:833-888) given (or drawing) the selection.
    Returns (loss, {"clip_ratio": ...}, sel)."""
    if sel is None:
        sel = cov_selection(kind, lp, olp, adv, cfg, mask, perm_fn)
    d = lp - olp
    ratio = d.exp()
    msl = getattr(cfg, "max_seq_len", None)
    if kind == "clip_cov":
        l1 = -adv * ratio
        l2 = -adv * ratio.clamp(1 - cfg.eps_clip_low, 1 + cfg.eps_clip_high)
        loss = reduce_loss(torch.maximum(l1, l2) * (1.0 - sel), mask, cfg.loss_reduction, msl)
        return loss, {"clip_ratio": masked_mean(sel, mask).item()}, sel
    tok = -adv * ratio + sel * cfg.kl_cov.ppo_kl_coef * d.abs()
    return reduce_loss(tok, mask, cfg.loss_reduction, msl), {"clip_ratio": 0.0}, sel


def loss_assembly(lp, olp, adv, cfg, mask=None, base=None, entropy=None):
    """workers/worker.py:852-889 — final = policy + kl*coef - entropy*coef; returns (final_loss, status)."""
    kind = cfg.policy_loss_type
    if kind in ("sapo", "cispo", "cross_entropy", "importance_sampling"):
        pl, met = extra_policy_loss(kind, lp, olp, adv, cfg, mask)
        clip_ratio = met.get("clip_ratio", 0.0)
        if kind == "cross_entropy":  # SFT path (worker.py:810-813): no KL / entropy term
            return pl, {"policy_loss": pl.item(), "final_loss": pl.item(), "loss_metrics/clip_ratio": 0.0}
    else:
        pl, clip_ratio = policy_loss(lp, olp, adv, cfg, mask)
    status = {"policy_loss": pl.item(), "loss_metrics/clip_ratio": clip_ratio}
    ent_term = torch.tensor(0.0)
    if entropy is not None:
        with torch.set_grad_enabled(bool(cfg.use_entropy_loss)):
            ent = masked_mean(entropy, mask)
        status["policy_entropy"] = ent.item()
        if cfg.use_entropy_loss:
            ent_term = ent * cfg.entropy_loss_coef
    kl = torch.tensor(0.0)
    if cfg.use_kl_loss and base is not None:
        kl = masked_mean(approx_kl(lp, base, mask, cfg.kl_estimator_type), mask, dim=-1).mean()
        status["policy_kl"] = kl.item()
    final = pl + kl * cfg.kl_loss_coef - ent_term
    status["final_loss"] = final.item()
    return final, status


# ---------------------------------------------------------------------------------------- logits / head
def logprobs_from_logits(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """utils/torch_utils.py:158-168 (fp32 branch): gather - logsumexp."""
    z = logits.float()
    return z.gather(-1, labels.unsqueeze(-1)).squeeze(-1) - torch.logsumexp(z, dim=-1)


def entropy_from_logits(logits: torch.Tensor, attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """utils/torch_utils.py:59-111: -(p * log p).sum(-1) [* mask]."""
    logp = F.log_softmax(logits.float(), dim=-1)
    ent = -(logp.exp() * logp).sum(-1)
    return ent * attention_mask if attention_mask is not None else ent


def head_logprobs_entropy(hidden_bf16, weight_bf16, labels, temperature: float = 1.0, chunk: int = 512):
    """model_wrapper.py:309-363 in fp32: logits = hidden @ W^T / T (fp32 math on the bf16-rounded inputs),
    then the two helpers above.  Differentiable w.r.t. hidden/weight when they are fp32 leaves."""
    h = hidden_bf16.float() if hidden_bf16.dtype != torch.float32 else hidden_bf16
    w = weight_bf16.float() if weight_bf16.dtype != torch.float32 else weight_bf16
    lps, ents = [], []
    for s in range(0, h.shape[0], chunk):
        z = (h[s:s + chunk] @ w.t()) / temperature
        lps.append(logprobs_from_logits(z, labels[s:s + chunk]))
        ents.append(entropy_from_logits(z))
    return torch.cat(lps), torch.cat(ents)


def make_cfg(**kw) -> SimpleNamespace:
    """AlgorithmConfig defaults (config/config.py:273-303) as a plain namespace."""
    base = dict(
        policy_loss_type="regular", loss_reduction="token_mean", eps_clip_low=0.2, eps_clip_high=0.2,
        clip_ratio_c=3.0, max_seq_len=None, use_kl_loss=True, kl_estimator_type="k3", kl_loss_coef=1e-3,
        use_entropy_loss=False, entropy_loss_coef=0.01, gamma=1.0, lambd=1.0, grpo_norm_by_std=True,
        off_policy_correction=SimpleNamespace(tis_ratio_type=None, sequence_mask_metric=None),
    )
    base.update(kw)
    return SimpleNamespace(**base)
