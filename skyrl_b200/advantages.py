"""Advantage estimators on B200 kernels, behind the reference's estimator signature.

Every function here is what ``AdvantageEstimatorRegistry.get(name)`` returns and is called exactly like the
reference calls its estimators (keywords only, skyrl-train/skyrl_train/utils/ppo_utils.py:1204-1214):

    fn(token_level_rewards=, response_mask=, index=, values=, grpo_norm_by_std=, gamma=, lambd=, config=, **kw)
        -> (advantages, returns)

Inputs may live on the CPU (the reference's driver passes CPU tensors, trainer.py:789/810) or on a CUDA device;
results come back on the device the rewards were on.  The arithmetic always runs in libskyb200 on the GPU —
there is no CPU path.
"""
from __future__ import annotations

import logging
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

logger = logging.getLogger("skyrl_b200")


def _device_for(t: torch.Tensor) -> torch.device:
    if t.is_cuda:
        return t.device
    if not torch.cuda.is_available():
        raise RuntimeError("skyrl_b200 needs a CUDA device: the advantage kernels have no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _f32(t: torch.Tensor, dev: torch.device) -> torch.Tensor:
    return t.detach().to(device=dev, dtype=torch.float32, non_blocking=True).contiguous()


def _mask_arg(mask: Optional[torch.Tensor], dev: torch.device):
    """(tensor_or_None, mask_type): fp32 and int64 masks are consumed as they are (bit-exact), others -> fp32."""
    if mask is None:
        return None, _lib.MASK_NONE
    m = mask.detach().to(device=dev, non_blocking=True)
    if m.dtype == torch.int64:
        return m.contiguous(), _lib.MASK_I64
    if m.dtype != torch.float32:
        m = m.to(torch.float32)
    return m.contiguous(), _lib.MASK_F32


def build_group_index(index: Sequence) -> Tuple[np.ndarray, np.ndarray, np.ndarray, int]:
    """uid list -> (group_of_row, group_start, group_rows, num_groups), all int32.

    Dense ids are assigned in order of first appearance; equality is Python ``==``/hash on the uid exactly
    as the reference's ``id2score[index[i]]`` dict lookup (ppo_utils.py:1164-1166), so group assignment is
    bit-exact for any hashable uid type (str, int, numpy scalar).
    """
    ids = {}
    group_of_row = np.empty(len(index), dtype=np.int32)
    for i in range(len(index)):
        key = index[i]
        if isinstance(key, np.generic):
            key = key.item()
        g = ids.get(key)
        if g is None:
            g = len(ids)
            ids[key] = g
        group_of_row[i] = g
    n_groups = len(ids)
    group_rows = np.argsort(group_of_row, kind="stable").astype(np.int32)
    counts = np.bincount(group_of_row, minlength=n_groups)
    group_start = np.zeros(n_groups + 1, dtype=np.int32)
    np.cumsum(counts, out=group_start[1:])
    return group_of_row, group_start, group_rows, n_groups


def _group_advantage(rewards: torch.Tensor, response_mask: torch.Tensor, index, mode: int,
                     epsilon: float, one_launch: Optional[bool] = None) -> torch.Tensor:
    """``one_launch``: None = the cluster form whenever the batch allows it, False = always the general two-kernel
    form, True = raise if the cluster form cannot take the batch (tests pin each path with it)."""
    lib = _lib.load()
    dev = _device_for(rewards)
    B, A = rewards.shape
    if len(index) != B:
        raise ValueError(f"index has {len(index)} entries for a batch of {B}")
    out_dev = rewards.device
    with torch.cuda.device(dev):
        r = _f32(rewards, dev)
        m, mtype = _mask_arg(response_mask, dev)
        gor, gstart, grows, ng = build_group_index(index)
        packed = torch.from_numpy(np.concatenate([gor, gstart, grows])).to(dev, non_blocking=True)
        g_of_row, g_start, g_rows = packed[:B], packed[B:B + ng + 1], packed[B + ng + 1:]
        out = torch.empty((B, A), dtype=torch.float32, device=dev)
        scores = torch.empty(B, dtype=torch.float32, device=dev)
        sizes = np.diff(gstart)
        n = int(sizes[0]) if ng > 0 else 0
        aligned = all(t is None or t.data_ptr() % 16 == 0 for t in (r, m, out))
        uniform = (ng > 0 and bool((sizes == n).all()) and aligned
                   and bool(lib.skyb200_group_advantage_uniform_supported(A, n)))
        if one_launch and not uniform:
            raise ValueError(f"the one-launch form needs uniform groups and a supported shape (A={A}, sizes={set(sizes.tolist())})")
        if uniform and one_launch is not False:
            # every prompt has n samples (the trainer's layout): one launch, a thread-block cluster per group
            _lib.check(
                lib.skyb200_group_advantage_uniform(_lib.ptr(r), _lib.ptr(m), mtype, _lib.ptr(g_rows), _lib.ptr(out),
                                                    _lib.ptr(scores), B, A, n, mode, float(epsilon),
                                                    _lib.stream_ptr()),
                "skyb200_group_advantage_uniform",
            )
        else:
            _lib.check(
                lib.skyb200_group_advantage(_lib.ptr(r), _lib.ptr(m), mtype, _lib.ptr(g_of_row), _lib.ptr(g_start),
                                            _lib.ptr(g_rows), _lib.ptr(out), _lib.ptr(scores), B, A, ng, mode,
                                            float(epsilon), _lib.stream_ptr()),
                "skyb200_group_advantage",
            )
        if mode == 2:
            if (sizes == 1).any():
                for g in np.nonzero(sizes == 1)[0]:
                    logger.warning("Only one response for prompt index %s, setting advantage to 0",
                                   index[int(grows[gstart[g]])])
    return out if out_dev == dev else out.to(out_dev)


def compute_grpo_outcome_advantage(
    token_level_rewards: torch.Tensor,
    response_mask: torch.Tensor,
    index,
    epsilon: float = 1e-6,
    grpo_norm_by_std: bool = True,
    **kwargs,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """GRPO outcome advantage (reference: ppo_utils.py:1132-1182).

    ``score_i = sum_t r[i, t]`` (unmasked); per uid group mean and *unbiased* std; ``(score - mean) / (std +
    eps)`` (mean 0 / std 1 for singleton groups); broadcast ``* response_mask``.  Advantages and returns are
    the same tensor, as in the reference.
    """
    adv = _group_advantage(token_level_rewards, response_mask, index, 0 if grpo_norm_by_std else 1, epsilon)
    return adv, adv


def compute_rloo_outcome_advantage(
    token_level_rewards: torch.Tensor,
    response_mask: torch.Tensor,
    index,
    **kwargs,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """RLOO advantage (reference: ppo_utils.py:1048-1098): ``(score - group_mean) * n / (n - 1)``, 0 if n == 1."""
    adv = _group_advantage(token_level_rewards, response_mask, index, 2, 0.0)
    return adv, adv


def compute_gae_advantage_return(
    token_level_rewards: torch.Tensor,
    values: torch.Tensor,
    response_mask: torch.Tensor,
    gamma: float,
    lambd: float,
    **kwargs,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """GAE + whitening (reference: ppo_utils.py:1101-1129, masked_whiten :163-172).

    The scan runs over all A positions of a row with no reset on the mask (the reference has none); returns
    are the un-whitened ``A_hat + V``; advantages are whitened with the global masked mean and Bessel-corrected
    variance and are *not* re-masked.  Raises ``ValueError`` when the mask sums to 0 or 1 (masked_var).
    """
    lib = _lib.load()
    if values is None:
        raise ValueError("GAE needs critic values")
    dev = _device_for(token_level_rewards)
    out_dev = token_level_rewards.device
    B, A = token_level_rewards.shape
    with torch.cuda.device(dev):
        r = _f32(token_level_rewards, dev)
        v = _f32(values, dev)
        m, mtype = _mask_arg(response_mask, dev)
        adv = torch.empty((B, A), dtype=torch.float32, device=dev)
        ret = torch.empty((B, A), dtype=torch.float32, device=dev)
        stats = torch.empty(4, dtype=torch.float32, device=dev)
        ws_bytes = lib.skyb200_gae_workspace(B)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(
            lib.skyb200_gae_advantage(_lib.ptr(r), _lib.ptr(v), _lib.ptr(m), mtype, _lib.ptr(adv), _lib.ptr(ret),
                                      _lib.ptr(stats), _lib.ptr(ws), ws_bytes, B, A, float(gamma), float(lambd),
                                      _lib.stream_ptr()),
            "skyb200_gae_advantage",
        )
        mask_sum = float(stats[0].item())  # the one sync of this call (the reference syncs on the same test)
    if mask_sum == 0:
        raise ValueError("At least one element in the mask has to be 1.")
    if mask_sum == 1:
        raise ValueError("The sum of the mask is one, which can cause a division by zero.")
    if out_dev != dev:
        adv, ret = adv.to(out_dev), ret.to(out_dev)
    return adv, ret


def compute_reinforce_plus_plus_outcome_advantage(
    token_level_rewards: torch.Tensor,
    response_mask: torch.Tensor,
    gamma: float,
    **kwargs,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """REINFORCE++ (reference: ppo_utils.py:1013-1045): discounted returns with a reset after EOS
    (``running *= mask[t]``), advantages = ``masked_whiten(returns, mask) * mask``."""
    lib = _lib.load()
    dev = _device_for(token_level_rewards)
    out_dev = token_level_rewards.device
    B, A = token_level_rewards.shape
    with torch.cuda.device(dev):
        r = _f32(token_level_rewards, dev)
        m, mtype = _mask_arg(response_mask, dev)
        adv = torch.empty((B, A), dtype=torch.float32, device=dev)
        ret = torch.empty((B, A), dtype=torch.float32, device=dev)
        stats = torch.empty(4, dtype=torch.float32, device=dev)
        ws_bytes = lib.skyb200_gae_workspace(B)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(
            lib.skyb200_reinforce_pp_advantage(_lib.ptr(r), _lib.ptr(m), mtype, _lib.ptr(adv), _lib.ptr(ret),
                                               _lib.ptr(stats), _lib.ptr(ws), ws_bytes, B, A, float(gamma),
                                               _lib.stream_ptr()),
            "skyb200_reinforce_pp_advantage",
        )
        mask_sum = float(stats[0].item())
    if mask_sum == 0:
        raise ValueError("At least one element in the mask has to be 1.")
    if mask_sum == 1:
        raise ValueError("The sum of the mask is one, which can cause a division by zero.")
    if out_dev != dev:
        adv, ret = adv.to(out_dev), ret.to(out_dev)
    return adv, ret


@torch.no_grad()
def normalize_advantages(advantages: torch.Tensor, response_mask: torch.Tensor) -> torch.Tensor:
    """Batch normalisation of advantages (reference: ``normalize_advantages_dict``, ppo_utils.py:127-145):
    mean over ALL positions, variance over the mask, ``(adv - mean) * rsqrt(clamp(var, 1e-8))``."""
    lib = _lib.load()
    dev = _device_for(advantages)
    out_dev = advantages.device
    with torch.cuda.device(dev):
        a = _f32(advantages, dev)
        m, mtype = _mask_arg(response_mask, dev)
        out = torch.empty_like(a)
        stats = torch.empty(4, dtype=torch.float32, device=dev)
        wsb = lib.skyb200_normalize_advantages_workspace()
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        _lib.check(
            lib.skyb200_normalize_advantages(_lib.ptr(a), _lib.ptr(m), mtype, _lib.ptr(out), _lib.ptr(stats),
                                             _lib.ptr(ws), wsb, a.numel(), _lib.stream_ptr()),
            "skyb200_normalize_advantages",
        )
    return out if out_dev == dev else out.to(out_dev)


def normalize_advantages_dict(data):
    """Dict / TrainingInputBatch form of :func:`normalize_advantages` (same keys as the reference)."""
    data["advantages"] = normalize_advantages(data["advantages"], data["response_mask"])
    return data


_KL_NAMES = ("k1", "abs", "k2", "k3")


def _kl_type(name: str) -> int:
    if name not in _lib.KL_TYPES:
        raise ValueError(f"Invalid KL estimator type: {name}")
    return _lib.KL_TYPES[name]


@torch.no_grad()
def compute_approx_kl(
    log_probs: torch.Tensor,
    log_probs_base: torch.Tensor,
    loss_mask: Optional[torch.Tensor] = None,
    kl_estimator_type: str = "k3",
) -> torch.Tensor:
    """k1 / abs / k2 / k3 KL estimators (reference: ppo_utils.py:88-124).  No gradient, like the reference."""
    lib = _lib.load()
    kt = _kl_type(kl_estimator_type)
    dev = _device_for(log_probs)
    out_dev = log_probs.device
    with torch.cuda.device(dev):
        lp, base = _f32(log_probs, dev), _f32(log_probs_base, dev)
        m = None if loss_mask is None else _f32(loss_mask, dev)
        out = torch.empty_like(lp)
        _lib.check(lib.skyb200_approx_kl(_lib.ptr(lp), _lib.ptr(base), _lib.ptr(m), _lib.ptr(out), lp.numel(), kt,
                                         _lib.stream_ptr()), "skyb200_approx_kl")
    return out if out_dev == dev else out.to(out_dev)


@torch.no_grad()
def apply_reward_kl_penalty(
    rewards: torch.Tensor,
    action_log_probs: torch.Tensor,
    base_action_log_probs: torch.Tensor,
    loss_mask: Optional[torch.Tensor],
    kl_estimator_type: str = "k3",
    kl_loss_coef: float = 1e-3,
):
    """KL-in-reward (reference: RayPPOTrainer.apply_reward_kl_penalty, trainer.py:981-1035).

    Returns ``(new_rewards, {"avg_kl", "avg_kl_max", "kl_loss_coef"})`` with
    ``new_rewards = rewards - kl * max(0, coef)``, ``avg_kl = mean_b masked_mean_t(kl)`` and
    ``avg_kl_max = mean_b max_t |kl|``.  The caller owns the KL controller update (trainer.py:1013-1015).
    """
    lib = _lib.load()
    kt = _kl_type(kl_estimator_type)
    dev = _device_for(rewards)
    out_dev = rewards.device
    B, A = rewards.shape
    with torch.cuda.device(dev):
        r, lp, base = _f32(rewards, dev), _f32(action_log_probs, dev), _f32(base_action_log_probs, dev)
        m = None if loss_mask is None else _f32(loss_mask, dev)
        out = torch.empty_like(r)
        metrics = torch.empty(2, dtype=torch.float32, device=dev)
        row_ws = torch.empty(2 * B, dtype=torch.float32, device=dev)
        _lib.check(
            lib.skyb200_reward_kl_penalty(_lib.ptr(lp), _lib.ptr(base), _lib.ptr(m), _lib.ptr(r), _lib.ptr(out),
                                          _lib.ptr(metrics), _lib.ptr(row_ws), B, A, kt, float(kl_loss_coef),
                                          _lib.stream_ptr()),
            "skyb200_reward_kl_penalty",
        )
        avg_kl, avg_kl_max = metrics.tolist()
    if out_dev != dev:
        out = out.to(out_dev)
    return out, {"avg_kl": avg_kl, "avg_kl_max": avg_kl_max, "kl_loss_coef": kl_loss_coef}


def compute_advantages_and_returns(
    token_level_rewards: torch.Tensor,
    response_mask: torch.Tensor,
    index,
    adv_estimator,
    config=None,
    values: Optional[torch.Tensor] = None,
    grpo_norm_by_std: bool = True,
    gamma=1.0,
    lambd=1.0,
    **kwargs,
):
    """Registry dispatch with the reference's keyword convention (ppo_utils.py:1190-1214)."""
    from .registry import AdvantageEstimatorRegistry

    fn = AdvantageEstimatorRegistry.get(adv_estimator)
    return fn(
        token_level_rewards=token_level_rewards,
        response_mask=response_mask,
        index=index,
        values=values,
        grpo_norm_by_std=grpo_norm_by_std,
        gamma=gamma,
        lambd=lambd,
        config=config,
        **kwargs,
    )
