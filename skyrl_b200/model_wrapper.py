"""The call-site glue of ``HFModelWrapper.forward`` (skyrl-train/skyrl_train/model_wrapper.py:261-375) for a trunk
that hands over hidden states instead of logits.

The reference computes log-probs and entropy for **all S positions** of every sequence and then slices
``[:, -num_actions-1:-1]`` (model_wrapper.py:370, worker.py:855).  Here only those A rows per sequence ever reach
the head: labels are ``roll(sequences, -1)`` (model_wrapper.py:291) sliced the same way, entropy is multiplied by
the attention mask of the same positions (model_wrapper.py:341-349), and the result has the reference's shapes.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .fused_head import fused_lm_head


def action_log_probs_from_hidden(
    hidden: torch.Tensor,
    lm_head_weight: torch.Tensor,
    sequences: torch.Tensor,
    num_actions: int,
    attention_mask: Optional[torch.Tensor] = None,
    temperature: float = 1.0,
    compute_entropy: bool = False,
    chunk: int = 0,
) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """``hidden`` [B, S, H] (last hidden state of the trunk), ``sequences`` [B, S] token ids.

    Returns ``(action_log_probs [B, A], entropy [B, A] or None)`` — exactly what the reference's wrapper returns as
    ``log_probs[:, -A-1:-1]`` and what the worker slices out of ``output["entropy"]``.
    """
    if hidden.dim() != 3 or sequences.shape != hidden.shape[:2]:
        raise ValueError(f"expected hidden [B, S, H] and sequences [B, S], got {tuple(hidden.shape)} {tuple(sequences.shape)}")
    B, S, H = hidden.shape
    A = int(num_actions)
    if not 0 < A < S:
        raise ValueError(f"num_actions must be in (0, {S}), got {A}")
    rolled = torch.roll(sequences, shifts=-1, dims=1)
    sl = slice(S - A - 1, S - 1)
    h = hidden[:, sl, :].reshape(B * A, H)
    labels = rolled[:, sl].reshape(B * A)
    emask = None
    if compute_entropy and attention_mask is not None:
        if attention_mask.shape != (B, S):
            raise ValueError(
                f"attention_mask shape {attention_mask.shape} does not match logits shape "
                f"(batch_size={B}, seqlen={S}). Expected attention_mask shape: ({B}, {S})"
            )
        emask = attention_mask[:, sl].reshape(B * A)
    lp, ent = fused_lm_head(h, lm_head_weight, labels, temperature=temperature, entropy_mask=emask,
                            compute_entropy=compute_entropy, chunk=chunk)
    return lp.view(B, A), (ent.view(B, A) if compute_entropy else None)
