"""Policy losses on the fused B200 loss kernel, behind the reference's ``PolicyLossRegistry`` signature.

    fn(log_probs, old_log_probs, advantages, config, loss_mask=None, rollout_logprobs=None)
        -> (loss: 0-d tensor with a grad path to log_probs, metrics: dict[str, float])

(reference: skyrl-train/skyrl_train/utils/ppo_utils.py:548-586 ``ppo_policy_loss`` for "regular"/"dual_clip",
:661-725 ``gspo_policy_loss``, :984-1009 ``reduce_loss``.)  One kernel pass computes the loss value, the
``clip_ratio`` metric and ``d loss / d log_probs``; backward only rescales by the upstream gradient.

:func:`policy_update_loss` additionally folds in the worker's loss assembly (worker.py:852-889): KL-loss term
(no gradient, like ``compute_approx_kl`` under ``@torch.no_grad``), entropy ``masked_mean`` and
``loss = policy + kl*coef - entropy*coef`` — still a single pass, one packed stats vector, one host sync.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _lib

_VALID_REDUCTIONS = ("token_mean", "sequence_mean", "seq_mean_token_sum_norm")
_ALL_REDUCTIONS = _VALID_REDUCTIONS + ("sum",)


def _cfg(config, name, default=None):
    """Attribute or mapping access (AlgorithmConfig dataclass, DictConfig, SimpleNamespace, dict)."""
    if isinstance(config, dict):
        return config.get(name, default)
    return getattr(config, name, default)


def _off_policy_enabled(config, rollout_logprobs) -> bool:
    if rollout_logprobs is None:
        return False
    opc = _cfg(config, "off_policy_correction", None)
    if opc is None:
        return False
    return _cfg(opc, "tis_ratio_type", None) is not None or _cfg(opc, "sequence_mask_metric", None) is not None


def _make_params(loss_type: str, reduction: str, config, *, kl_type: str = "k3", kl_coef: float = 0.0,
                 use_entropy_loss: bool = False, entropy_coef: float = 0.0,
                 kl_has_grad: bool = False) -> _lib.LossParams:
    if reduction not in _ALL_REDUCTIONS:
        raise ValueError(f"Invalid loss reduction type: {reduction}")
    max_seq_len = _cfg(config, "max_seq_len", None)
    if reduction == "seq_mean_token_sum_norm":
        assert max_seq_len is not None, "max_seq_len must be provided for seq_mean_token_sum_norm loss reduction"
    if kl_type not in _lib.KL_TYPES:
        raise ValueError(f"Invalid KL estimator type: {kl_type}")
    p = _lib.LossParams()
    p.loss_type = _lib.LOSS_TYPES[loss_type]
    p.reduction = _lib.REDUCTIONS[reduction]
    p.kl_type = _lib.KL_TYPES[kl_type]
    p.kl_has_grad = int(kl_has_grad)
    p.use_entropy_loss = int(use_entropy_loss)
    if loss_type == "cispo":  # CISPO clips with its own epsilons (config.cispo.*, ppo_utils.py:746)
        cis = _cfg(config, "cispo", None)
        p.eps_clip_low = float(_cfg(cis, "cispo_eps_clip_low"))
        p.eps_clip_high = float(_cfg(cis, "cispo_eps_clip_high"))
    else:
        p.eps_clip_low = float(_cfg(config, "eps_clip_low", 0.2))
        p.eps_clip_high = float(_cfg(config, "eps_clip_high", 0.2))
    if loss_type == "sapo":
        sapo = _cfg(config, "sapo", None)
        p.sapo_tau_pos = float(_cfg(sapo, "tau_pos"))
        p.sapo_tau_neg = float(_cfg(sapo, "tau_neg"))
    p.clip_ratio_c = float(_cfg(config, "clip_ratio_c", 3.0))
    p.max_seq_len = float(max_seq_len) if max_seq_len is not None else 0.0
    p.kl_loss_coef = float(kl_coef)
    p.entropy_loss_coef = float(entropy_coef)
    return p


def _prep(t: Optional[torch.Tensor], ref: torch.Tensor) -> Optional[torch.Tensor]:
    if t is None:
        return None
    t = t.detach()
    if t.device != ref.device or t.dtype != torch.float32:
        t = t.to(device=ref.device, dtype=torch.float32)
    return t.contiguous()


def _run_kernel(lp, olp, adv, mask, base, ent, params, want_d_ent, tis=None, clip_mask=None):
    """Launch skyb200_policy_loss. Returns (stats[16] device tensor, d_lp, d_ent or None)."""
    lib = _lib.load()
    B, A = lp.shape
    with torch.cuda.device(lp.device):
        stats = torch.empty(_lib.LOSS_NSTATS, dtype=torch.float32, device=lp.device)
        d_lp = torch.empty_like(lp)
        d_ent = torch.empty_like(lp) if want_d_ent else None
        ws_bytes = lib.skyb200_policy_loss_workspace(B, A)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=lp.device)
        _lib.check(
            lib.skyb200_policy_loss(_lib.ptr(lp), _lib.ptr(olp), _lib.ptr(adv), _lib.ptr(mask), _lib.ptr(base),
                                    _lib.ptr(ent), _lib.ptr(d_lp), _lib.ptr(d_ent), _lib.ptr(stats), _lib.ptr(ws),
                                    ws_bytes, B, A, C.byref(params), _lib.ptr(tis), _lib.ptr(clip_mask),
                                    _lib.stream_ptr()),
            "skyb200_policy_loss",
        )
    return stats, d_lp, d_ent


def _scale_by_upstream(grad: torch.Tensor, upstream: torch.Tensor) -> torch.Tensor:
    """grad *= upstream (0-d device tensor) without a host sync."""
    lib = _lib.load()
    up = upstream.detach().to(device=grad.device, dtype=torch.float32).reshape(1).contiguous()
    with torch.cuda.device(grad.device):
        _lib.check(lib.skyb200_scale_inplace(_lib.ptr(grad), _lib.ptr(up), grad.numel(), _lib.stream_ptr()),
                   "skyb200_scale_inplace")
    return grad


class _PolicyLossFn(torch.autograd.Function):
    """loss, stats = f(log_probs [, entropy]); gradients were produced by the forward kernel."""

    @staticmethod
    def forward(ctx, log_probs, entropy, old_log_probs, advantages, loss_mask, base_log_probs, params, stat_index,
                tis=None, clip_mask=None):
        if not log_probs.is_cuda:
            raise RuntimeError("skyrl_b200 policy losses need CUDA tensors (no CPU fallback)")
        if log_probs.dim() != 2:
            raise ValueError(f"log_probs must be [batch, actions], got {tuple(log_probs.shape)}")
        lp = _prep(log_probs, log_probs)
        olp = _prep(old_log_probs, lp)
        adv = _prep(advantages, lp)
        mask = _prep(loss_mask, lp)
        base = _prep(base_log_probs, lp)
        ent = _prep(entropy, lp)
        want_d_ent = ent is not None and entropy.requires_grad and bool(params.use_entropy_loss)
        stats, d_lp, d_ent = _run_kernel(lp, olp, adv, mask, base, ent, params, want_d_ent, _prep(tis, lp),
                                         _prep(clip_mask, lp))
        ctx.save_for_backward(d_lp, d_ent if d_ent is not None else torch.empty(0, device=lp.device))
        ctx.has_d_ent = d_ent is not None
        ctx.in_dtype = log_probs.dtype
        ctx.ent_dtype = entropy.dtype if entropy is not None else None
        ctx.mark_non_differentiable(stats)
        loss = stats[stat_index].clone()
        return loss, stats

    @staticmethod
    def backward(ctx, g_loss, _g_stats):
        d_lp, d_ent = ctx.saved_tensors
        g_lp = _scale_by_upstream(d_lp.clone(), g_loss)
        if g_lp.dtype != ctx.in_dtype:
            g_lp = g_lp.to(ctx.in_dtype)
        g_ent = None
        if ctx.has_d_ent:
            g_ent = _scale_by_upstream(d_ent.clone(), g_loss)
            if g_ent.dtype != ctx.ent_dtype:
                g_ent = g_ent.to(ctx.ent_dtype)
        return g_lp, g_ent, None, None, None, None, None, None, None, None


def compute_off_policy_correction(old_log_probs, rollout_logprobs, loss_mask, off_policy_correction):
    """TIS weights, outlier-token mask and sequence mask (reference: off_policy_correction_utils.py:173-262).

    Returns ``(tis_ratio [B, A] or None, metrics dict, new_loss_mask [B, A])`` with the reference's metric names.
    """
    opc = off_policy_correction
    tis_type = _cfg(opc, "tis_ratio_type", None)
    seq_metric = _cfg(opc, "sequence_mask_metric", None)
    if tis_type is None and seq_metric is None:
        return None, {}, loss_mask
    if tis_type not in (None, "token", "sequence"):
        raise ValueError(f"Unknown tis_ratio_type: {tis_type}")
    if seq_metric not in (None, "geometric", "product"):
        raise ValueError(f"Unknown sequence_mask_metric: {seq_metric}")
    lib = _lib.load()
    if not old_log_probs.is_cuda:
        raise RuntimeError("skyrl_b200 off-policy correction needs CUDA tensors (no CPU fallback)")
    olp = _prep(old_log_probs, old_log_probs)
    rlp = _prep(rollout_logprobs, olp)
    mask = _prep(loss_mask, olp) if loss_mask is not None else torch.ones_like(olp)
    B, A = olp.shape
    p = _lib.OpcParams()
    p.tis_type = {None: 0, "token": 1, "sequence": 2}[tis_type]
    p.seq_mask_type = {None: 0, "geometric": 1, "product": 2}[seq_metric]
    lo, hi = _cfg(opc, "outlier_token_is_threshold_low", None), _cfg(opc, "outlier_token_is_threshold_high", None)
    p.has_outlier_low, p.has_outlier_high = int(lo is not None), int(hi is not None)
    p.outlier_low, p.outlier_high = float(lo or 0.0), float(hi or 0.0)
    p.token_tis_cap = float(_cfg(opc, "token_tis_ratio_clip_high", 2.0))
    p.seq_tis_cap = float(_cfg(opc, "sequence_tis_ratio_clip_high", 5.0))
    p.geo_low, p.geo_high = float(_cfg(opc, "geo_mask_low", 0.99)), float(_cfg(opc, "geo_mask_high", 1.01))
    p.prod_low, p.prod_high = float(_cfg(opc, "product_mask_low", 0.5)), float(_cfg(opc, "product_mask_high", 2.0))
    with torch.cuda.device(olp.device):
        tis = torch.empty_like(olp)
        mask_out = torch.empty_like(olp)
        met = torch.empty(_lib.OPC_NMETRICS, dtype=torch.float32, device=olp.device)
        wsb = lib.skyb200_off_policy_workspace(B)
        ws = torch.empty(wsb, dtype=torch.uint8, device=olp.device)
        _lib.check(
            lib.skyb200_off_policy_correction(_lib.ptr(olp), _lib.ptr(rlp), _lib.ptr(mask), _lib.ptr(tis),
                                              _lib.ptr(mask_out), _lib.ptr(met), _lib.ptr(ws), wsb, B, A, C.byref(p),
                                              _lib.stream_ptr()),
            "skyb200_off_policy_correction",
        )
        m = met.tolist()
    metrics = {"is_ratio_mean": m[0], "is_ratio_std": m[1], "is_ratio_max": m[2], "is_ratio_min": m[3],
               "outlier_seq_masked_ratio": m[4], "outlier_seq_over_high_ratio": m[5],
               "outlier_seq_under_low_ratio": m[6]}
    if tis_type == "token":
        metrics["tis_token_clip_high_ratio"] = m[7]
    elif tis_type == "sequence":
        metrics["tis_seq_clip_high_ratio"] = m[7]
    if seq_metric is not None:
        name = "geo" if seq_metric == "geometric" else "product"
        metrics[f"{name}_sequence_mask_masked_ratio"] = m[8]
        metrics[f"{name}_sequence_mask_over_high_ratio"] = m[9]
        metrics[f"{name}_sequence_mask_under_low_ratio"] = m[10]
    return (tis if tis_type is not None else None), metrics, mask_out


def _registry_loss(loss_type, log_probs, old_log_probs, advantages, config, loss_mask, rollout_logprobs,
                   reduction=None, want_stats=False):
    tis, opc_metrics, clip_mask = None, {}, None
    if loss_type not in ("cross_entropy", "importance_sampling") and _off_policy_enabled(config, rollout_logprobs):
        # the clip_ratio metric keeps the mask it had before the correction (ppo_utils.py:570 vs :579-583)
        clip_mask = loss_mask if loss_mask is not None else torch.ones_like(old_log_probs, dtype=torch.float32)
        tis, opc_metrics, loss_mask = compute_off_policy_correction(
            old_log_probs, rollout_logprobs, loss_mask, _cfg(config, "off_policy_correction"))
    params = _make_params(loss_type, reduction or _cfg(config, "loss_reduction"), config)
    loss, stats = _PolicyLossFn.apply(log_probs, None, old_log_probs, advantages, loss_mask, None, params,
                                      _lib.STAT_POLICY_LOSS, tis, clip_mask)
    if want_stats:
        return loss, stats
    clip_ratio = float(stats[_lib.STAT_CLIP_RATIO].item())  # the single host sync of this call
    metrics = {"clip_ratio": clip_ratio}
    metrics.update(opc_metrics)
    return loss, metrics


def ppo_policy_loss(
    log_probs: torch.Tensor,
    old_log_probs: torch.Tensor,
    advantages: torch.Tensor,
    config,
    loss_mask: Optional[torch.Tensor] = None,
    rollout_logprobs: Optional[torch.Tensor] = None,
) -> Tuple[torch.Tensor, Dict[str, float]]:
    """"regular" and "dual_clip" PPO losses (reference: ppo_utils.py:548-586)."""
    loss_type = _cfg(config, "policy_loss_type")
    assert loss_type in ["regular", "dual_clip"], "loss_type must be either 'regular' or 'dual_clip'"
    assert _cfg(config, "loss_reduction") in _VALID_REDUCTIONS, (
        "loss_reduction must be either 'token_mean', 'sequence_mean', or 'seq_mean_token_sum_norm'"
    )
    return _registry_loss(loss_type, log_probs, old_log_probs, advantages, config, loss_mask, rollout_logprobs)


def gspo_policy_loss(
    log_probs: torch.Tensor,
    old_log_probs: torch.Tensor,
    advantages: torch.Tensor,
    config,
    loss_mask: Optional[torch.Tensor] = None,
    rollout_logprobs: Optional[torch.Tensor] = None,
) -> Tuple[torch.Tensor, Dict[str, float]]:
    """GSPO sequence-level importance ratio loss (reference: ppo_utils.py:661-725)."""
    if _cfg(config, "loss_reduction") != "sequence_mean":
        import logging

        logging.getLogger("skyrl_b200").warning(
            "With GSPO it's recommended to use 'sequence_mean' loss reduction; got %s", _cfg(config, "loss_reduction")
        )
    return _registry_loss("gspo", log_probs, old_log_probs, advantages, config, loss_mask, rollout_logprobs)


def sapo_policy_loss(log_probs, old_log_probs, advantages, config, loss_mask=None, rollout_logprobs=None):
    """SAPO soft-gated objective (reference: ppo_utils.py:589-658); ``clip_ratio`` is reported as 0.0."""
    if _cfg(config, "loss_reduction") != "sequence_mean":
        import logging

        logging.getLogger("skyrl_b200").warning(
            "With SAPO it's recommended to use 'sequence_mean' loss reduction; got %s", _cfg(config, "loss_reduction")
        )
    loss, met = _registry_loss("sapo", log_probs, old_log_probs, advantages, config, loss_mask, rollout_logprobs)
    met["clip_ratio"] = 0.0
    return loss, met


def compute_policy_loss_cispo(log_probs, old_log_probs, advantages, config, loss_mask=None, rollout_logprobs=None):
    """CISPO: clipped, detached importance weight times log-prob (reference: ppo_utils.py:728-761)."""
    return _registry_loss("cispo", log_probs, old_log_probs, advantages, config, loss_mask, rollout_logprobs)


def cross_entropy_loss(log_probs, old_log_probs, advantages, config, loss_mask=None, rollout_logprobs=None):
    """SFT negative log-likelihood, masked SUM (Tinker semantics; reference: ppo_utils.py:892-930).

    ``old_log_probs`` and ``advantages`` are ignored, as in the reference, and may be ``None``."""
    if old_log_probs is None:  # the kernel signature wants the arrays; SFT never reads their values
        old_log_probs = log_probs.detach()
    if advantages is None:
        advantages = log_probs.detach()
    loss, _ = _registry_loss("cross_entropy", log_probs, old_log_probs, advantages, config, loss_mask, None,
                             reduction="sum", want_stats=True)
    return loss, {"clip_ratio": 0.0}


def importance_sampling_loss(log_probs, old_log_probs, advantages, config, loss_mask=None, rollout_logprobs=None):
    """``-(exp(lp - old_lp) * A)`` masked SUM (Tinker semantics; reference: ppo_utils.py:933-981)."""
    loss, stats = _registry_loss("importance_sampling", log_probs, old_log_probs, advantages, config, loss_mask, None,
                                 reduction="sum", want_stats=True)
    return loss, {"importance_ratio": float(stats[_lib.STAT_RATIO_MEAN].item())}


def cov_token_selection(kind, log_probs, old_log_probs, advantages, config, loss_mask=None, *, value_selected=1.0,
                        value_other=0.0, seed=None):
    """The detached token selection of ``clip_cov`` (ppo_utils.py:785-814) / ``kl_cov`` (:857-871) as an fp32
    ``[B, A]`` tensor (``value_selected`` on the chosen tokens), plus a device tensor ``[n_selected, n_candidates]``.

    ``kl_cov`` is deterministic (top-k of the covariance).  ``clip_cov`` draws its random subset with a counter-based
    generator on the device; ``seed`` defaults to a draw from torch's global CPU generator, so ``torch.manual_seed``
    makes it reproducible like the reference's ``torch.randperm`` (the subset itself differs from torch's)."""
    if kind not in ("clip_cov", "kl_cov"):
        raise ValueError(kind)
    if not log_probs.is_cuda:
        raise RuntimeError("skyrl_b200 covariance losses need CUDA tensors (no CPU fallback)")
    lib = _lib.load()
    lp = _prep(log_probs, log_probs)
    olp, adv, mask = _prep(old_log_probs, lp), _prep(advantages, lp), _prep(loss_mask, lp)
    B, A = lp.shape
    p = _lib.CovParams()
    p.value_selected, p.value_other = float(value_selected), float(value_other)
    if kind == "clip_cov":
        c = _cfg(config, "clip_cov")
        p.mode = _lib.COV_CLIP
        p.eps_clip_low, p.eps_clip_high = float(_cfg(config, "eps_clip_low", 0.2)), float(_cfg(config, "eps_clip_high", 0.2))
        p.cov_lb, p.cov_ub = float(_cfg(c, "clip_cov_lb")), float(_cfg(c, "clip_cov_ub"))
        p.frac = float(_cfg(c, "clip_ratio"))
        p.seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if seed is None else int(seed)
    else:
        p.mode = _lib.COV_KL
        p.frac = float(_cfg(_cfg(config, "kl_cov"), "kl_cov_frac"))
    with torch.cuda.device(lp.device):
        out = torch.empty_like(lp)
        counts = torch.zeros(2, dtype=torch.float32, device=lp.device)
        wsb = lib.skyb200_cov_select_workspace(B, A)
        ws = torch.empty(wsb, dtype=torch.uint8, device=lp.device)
        _lib.check(lib.skyb200_cov_select(_lib.ptr(lp), _lib.ptr(olp), _lib.ptr(adv), _lib.ptr(mask), _lib.ptr(out),
                                          _lib.ptr(counts), _lib.ptr(ws), wsb, B, A, C.byref(p), _lib.stream_ptr()),
                   "skyb200_cov_select")
    return out, counts


def compute_policy_loss_clip_cov(log_probs, old_log_probs, advantages, config, loss_mask=None, rollout_logprobs=None):
    """Clip-Cov (reference: ppo_utils.py:765-829): the PPO clipped loss with a random subset of the
    covariance-candidate tokens dropped; ``clip_ratio`` = masked fraction of dropped tokens."""
    if loss_mask is None:  # the reference indexes with the mask unconditionally
        loss_mask = torch.ones_like(advantages, dtype=torch.float32)
    corr, _ = cov_token_selection("clip_cov", log_probs.detach(), old_log_probs, advantages, config, loss_mask,
                                  value_selected=0.0, value_other=1.0)
    params = _make_params("regular", _cfg(config, "loss_reduction"), config)
    loss, _stats = _PolicyLossFn.apply(log_probs, None, old_log_probs, advantages, loss_mask, None, params,
                                       _lib.STAT_POLICY_LOSS, corr, None)  # corr rides in as the detached weight
    from .torch_utils import masked_mean

    clip_frac = masked_mean(1.0 - corr, _prep(loss_mask, corr))
    return loss, {"clip_ratio": float(clip_frac.item())}


def compute_policy_loss_kl_cov(log_probs, old_log_probs, advantages, config, loss_mask=None, rollout_logprobs=None):
    """KL-Cov (reference: ppo_utils.py:833-888): ``-A * ratio`` plus ``ppo_kl_coef * |lp - old_lp|`` on the
    ``kl_cov_frac`` of valid tokens with the largest covariance; ``clip_ratio`` is reported as 0.0."""
    if loss_mask is None:
        loss_mask = torch.ones_like(advantages, dtype=torch.float32)
    coef = float(_cfg(_cfg(config, "kl_cov"), "ppo_kl_coef"))
    w, _ = cov_token_selection("kl_cov", log_probs.detach(), old_log_probs, advantages, config, loss_mask,
                               value_selected=coef, value_other=0.0)
    params = _make_params("regular", _cfg(config, "loss_reduction"), config)
    params.loss_type = _lib.LOSS_KL_COV
    loss, _stats = _PolicyLossFn.apply(log_probs, None, old_log_probs, advantages, loss_mask, None, params,
                                       _lib.STAT_POLICY_LOSS, w, None)
    return loss, {"clip_ratio": 0.0}


class _CriticLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, values, old_values, returns, loss_mask, value_clip):
        if not values.is_cuda:
            raise RuntimeError("skyrl_b200 critic loss needs CUDA tensors (no CPU fallback)")
        lib = _lib.load()
        v = _prep(values, values)
        ov = _prep(old_values, v)
        ret = _prep(returns, v)
        mask = _prep(loss_mask, v)
        B, A = v.shape
        with torch.cuda.device(v.device):
            stats = torch.empty(4, dtype=torch.float32, device=v.device)
            d_v = torch.empty_like(v)
            ws_bytes = lib.skyb200_policy_loss_workspace(B, A)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=v.device)
            use_clip = value_clip is not None
            _lib.check(
                lib.skyb200_critic_loss(_lib.ptr(v), _lib.ptr(ov), _lib.ptr(ret), _lib.ptr(mask), _lib.ptr(d_v),
                                        _lib.ptr(stats), _lib.ptr(ws), ws_bytes, B, A,
                                        float(value_clip) if use_clip else 0.0, int(use_clip), _lib.stream_ptr()),
                "skyb200_critic_loss",
            )
        ctx.save_for_backward(d_v)
        ctx.in_dtype = values.dtype
        ctx.mark_non_differentiable(stats)
        return stats[0].clone(), stats

    @staticmethod
    def backward(ctx, g_loss, _g_stats):
        (d_v,) = ctx.saved_tensors
        g = _scale_by_upstream(d_v.clone(), g_loss)
        return (g if g.dtype == ctx.in_dtype else g.to(ctx.in_dtype)), None, None, None, None


def ppo_critic_loss(
    values: torch.Tensor,
    old_values: torch.Tensor,
    returns: torch.Tensor,
    config,
    loss_mask: Optional[torch.Tensor] = None,
):
    """Clipped value loss (reference: ppo_utils.py:175-193). Returns ``(0.5 * loss, clipfrac or None)``."""
    value_clip = _cfg(config, "value_clip", None)
    loss, stats = _CriticLossFn.apply(values, old_values, returns, loss_mask, value_clip)
    clipfrac = float(stats[1].item()) if value_clip is not None else None
    return loss, clipfrac


def policy_update_loss(
    log_probs: torch.Tensor,
    old_log_probs: torch.Tensor,
    advantages: torch.Tensor,
    config,
    loss_mask: Optional[torch.Tensor] = None,
    base_log_probs: Optional[torch.Tensor] = None,
    entropy: Optional[torch.Tensor] = None,
    policy_loss_type: Optional[str] = None,
    kl_has_grad: bool = False,
    sync: bool = True,
):
    """The worker's whole loss assembly in one kernel pass (reference: worker.py:787-889).

    ``loss = policy_loss + kl_loss * kl_loss_coef - entropy * entropy_loss_coef`` with
    ``kl_loss = mean_b masked_mean_t compute_approx_kl(...)`` (iff ``use_kl_loss``; contributes to the value,
    not to the gradient, exactly like the reference) and ``entropy = masked_mean(entropy, loss_mask)`` (gradient
    iff ``use_entropy_loss``).

    Returns ``(final_loss, status)``; ``status`` has the reference's metric names (``final_loss``,
    ``policy_loss``, ``policy_entropy``, ``policy_kl``, ``loss_metrics/clip_ratio``).  With ``sync=False`` the
    status values stay on the device as a packed fp32 vector (key ``"packed"``) for a later single all-reduce.
    """
    loss_type = policy_loss_type or _cfg(config, "policy_loss_type", "regular")
    if loss_type not in _lib.LOSS_TYPES:
        raise ValueError(f"policy loss '{loss_type}' has no fused B200 kernel; available: {list(_lib.LOSS_TYPES)}")
    use_kl = bool(_cfg(config, "use_kl_loss", False)) and base_log_probs is not None
    use_ent = bool(_cfg(config, "use_entropy_loss", False))
    reduction = _cfg(config, "loss_reduction", "token_mean")
    if loss_type in ("cross_entropy", "importance_sampling"):
        reduction = "sum"  # Tinker semantics: masked SUM whatever loss_reduction says (ppo_utils.py:892-981)
    if loss_type == "cross_entropy":
        use_kl = use_ent = False  # the SFT path of the worker adds no KL / entropy term (worker.py:810-813)
    params = _make_params(
        loss_type,
        reduction,
        config,
        kl_type=_cfg(config, "kl_estimator_type", "k3"),
        kl_coef=float(_cfg(config, "kl_loss_coef", 0.0)) if use_kl else 0.0,
        use_entropy_loss=use_ent,
        entropy_coef=float(_cfg(config, "entropy_loss_coef", 0.0)),
        kl_has_grad=kl_has_grad,
    )
    loss, stats = _PolicyLossFn.apply(log_probs, entropy, old_log_probs, advantages, loss_mask,
                                      base_log_probs if use_kl else None, params, _lib.STAT_FINAL_LOSS)
    if not sync:
        return loss, {"packed": stats}
    s = stats.tolist()  # one D2H copy for every metric
    status = {
        "final_loss": s[_lib.STAT_FINAL_LOSS],
        "policy_loss": s[_lib.STAT_POLICY_LOSS],
        "policy_entropy": s[_lib.STAT_ENTROPY],
        "loss_metrics/clip_ratio": s[_lib.STAT_CLIP_RATIO],
    }
    if use_kl:
        status["policy_kl"] = s[_lib.STAT_KL_LOSS]
    return loss, status


def status_from_packed(stats_host, use_kl: bool) -> Dict[str, float]:
    """Decode a packed stats vector (list of floats) into the reference's status dict."""
    status = {
        "final_loss": stats_host[_lib.STAT_FINAL_LOSS],
        "policy_loss": stats_host[_lib.STAT_POLICY_LOSS],
        "policy_entropy": stats_host[_lib.STAT_ENTROPY],
        "loss_metrics/clip_ratio": stats_host[_lib.STAT_CLIP_RATIO],
    }
    if use_kl:
        status["policy_kl"] = stats_host[_lib.STAT_KL_LOSS]
    return status
