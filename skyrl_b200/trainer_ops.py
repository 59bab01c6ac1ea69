"""Driver-side steps of the reference trainer that sit on the hot path, on the B200 kernels.

* :func:`compute_advantages_and_returns_batch` — ``RayPPOTrainer.compute_advantages_and_returns``
  (skyrl-train/skyrl_train/trainer.py:760-862), including the step-wise trajectory variant (:777-808) and the
  four metrics it logs.
* :func:`apply_reward_kl_penalty_batch` — ``RayPPOTrainer.apply_reward_kl_penalty`` (trainer.py:981-1035) including
  the KL-controller update.

``data`` is any mapping with the reference's keys (a ``TrainingInputBatch`` or a plain dict of tensors); ``uids`` are
passed explicitly because the reference keeps them in ``data.metadata``.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .advantages import _device_for, _f32, apply_reward_kl_penalty, compute_advantages_and_returns


def gather_rows(src: torch.Tensor, row_index: torch.Tensor) -> torch.Tensor:
    """``src[row_index]`` for fp32 ``[T, A]`` on the GPU (one coalesced pass)."""
    lib = _lib.load()
    dev = _device_for(src)
    out_dev = src.device
    with torch.cuda.device(dev):
        s = _f32(src, dev)
        idx = row_index.to(device=dev, dtype=torch.int32).contiguous()
        B, A = idx.numel(), s.shape[1]
        out = torch.empty((B, A), dtype=torch.float32, device=dev)
        _lib.check(lib.skyb200_gather_rows(_lib.ptr(s), _lib.ptr(idx), _lib.ptr(out), s.shape[0], B, A,
                                           _lib.stream_ptr()), "skyb200_gather_rows")
    return out if out_dev == dev else out.to(out_dev)


def compute_advantages_and_returns_batch(
    data,
    uids: Sequence,
    algorithm_cfg,
    step_wise_trajectories: bool = False,
    pad_size: int = 0,
) -> Tuple[Dict[str, torch.Tensor], Dict[str, float]]:
    """Adds ``data["advantages"]`` / ``data["returns"]`` and returns ``(data, metrics)`` with the reference's
    ``avg_final_rewards``, ``avg_advantages``, ``avg_advantages_abs`` (trainer.py:760-862)."""
    from .losses import _cfg

    rewards = data["rewards"]
    kw = dict(
        adv_estimator=_cfg(algorithm_cfg, "advantage_estimator", "grpo"),
        config=algorithm_cfg,
        gamma=_cfg(algorithm_cfg, "gamma", 1.0),
        lambd=_cfg(algorithm_cfg, "lambd", 1.0),
        grpo_norm_by_std=_cfg(algorithm_cfg, "grpo_norm_by_std", True),
    )
    values = data.get("values") if hasattr(data, "get") else data["values"]
    index = np.asarray(uids)
    if step_wise_trajectories:
        is_last = data["is_last_step"].bool()
        sel = is_last.cpu().numpy()
        adv_last, ret_last = compute_advantages_and_returns(
            token_level_rewards=rewards[is_last], response_mask=data["response_mask"][is_last], index=index[sel],
            values=values[is_last] if values is not None else None, **kw)
        # trajectory id of every step row = number of completed trajectories before it (trainer.py:795-797)
        traj_ids = torch.cat([torch.zeros(1, dtype=torch.bool, device=is_last.device), is_last[:-1]]).int().cumsum(0)
        num_groups = int(traj_ids[-1].item()) + 1
        assert num_groups == len(adv_last), (
            f"number of groups {num_groups} doesn't match the number of trajectories as given by `is_last_step` "
            f"{len(adv_last)}. The `is_last_step` tensor is likely malformed"
        )
        advantages = gather_rows(adv_last, traj_ids)
        returns = advantages if ret_last is adv_last else gather_rows(ret_last, traj_ids)
    else:
        advantages, returns = compute_advantages_and_returns(
            token_level_rewards=rewards, response_mask=data["response_mask"], index=index, values=values, **kw)
    data["advantages"], data["returns"] = advantages, returns

    n = len(rewards) - pad_size
    return_sums = rewards.sum(dim=-1)[:n]
    if step_wise_trajectories:
        avg_rewards = return_sums[data["is_last_step"][:n].bool()].mean().item()
    else:
        avg_rewards = return_sums.mean().item()
    mask = data["response_mask"][:n].to(advantages.device)
    adv = advantages[:n]
    from .torch_utils import masked_mean

    m = mask.to(torch.float32)
    if adv.is_cuda:
        avg_adv = masked_mean(adv.contiguous(), m.contiguous()).item()
        avg_abs = masked_mean(adv.abs().contiguous(), m.contiguous()).item()
    else:
        valid = torch.masked_select(adv, mask.bool())
        avg_adv, avg_abs = valid.mean().item(), valid.abs().mean().item()
    return data, {"avg_final_rewards": avg_rewards, "avg_advantages": avg_adv, "avg_advantages_abs": avg_abs}


def apply_reward_kl_penalty_batch(data, algorithm_cfg, reward_kl_controller=None):
    """``rewards -= kl * max(0, coef)`` + metrics + controller update (trainer.py:981-1035)."""
    from .losses import _cfg

    coef = reward_kl_controller.value if reward_kl_controller is not None else _cfg(algorithm_cfg, "kl_loss_coef")
    new_rewards, met = apply_reward_kl_penalty(data["rewards"], data["action_log_probs"],
                                               data["base_action_log_probs"], data["loss_mask"],
                                               _cfg(algorithm_cfg, "kl_estimator_type", "k3"), coef)
    data["rewards"] = new_rewards
    if reward_kl_controller is not None:
        reward_kl_controller.update(current=met["avg_kl"], n_steps=data["rewards"].shape[0])
    return data, met


class AdaptiveKLController:
    """Adaptive KL coefficient (reference: ppo_utils.py:45-60, https://arxiv.org/pdf/1909.08593.pdf)."""

    def __init__(self, init_kl_coef, target, horizon):
        self.value, self.target, self.horizon = init_kl_coef, target, horizon

    def update(self, current, n_steps):
        err = float(np.clip(current / self.target - 1, -0.2, 0.2))
        self.value *= 1 + err * n_steps / self.horizon


class FixedKLController:
    """Fixed KL coefficient (reference: ppo_utils.py:63-70)."""

    def __init__(self, kl_coef):
        self.value = kl_coef

    def update(self, current, n_steps):
        pass
