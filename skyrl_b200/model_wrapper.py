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


def action_log_probs_from_packed_hidden(
    hidden_packed: torch.Tensor,
    lm_head_weight: torch.Tensor,
    sequences: torch.Tensor,
    attention_mask: torch.Tensor,
    num_actions: int,
    temperature: float = 1.0,
    compute_entropy: bool = False,
    chunk: int = 0,
) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """The sample-packing branch of the wrapper (model_wrapper.py:278-289, 324-336, 353-358): the trunk ran on the
    un-padded token stream, ``hidden_packed`` is its last hidden state ``[nnz, H]`` (or ``[1, nnz, H]``) in
    ``unpad_input`` order (row-major over the valid positions of ``attention_mask`` [B, S]).

    The reference projects all ``nnz`` tokens, pads the log-probs back to ``[B, S]`` with zeros and slices
    ``[:, -A-1:-1]``.  Here only the packed rows that land inside that window reach the fused head; labels are the
    packed stream rolled by one (``model_wrapper.py:291`` — including the reference's documented quirk that the last
    token of sequence n is scored against the first token of sequence n+1), masked positions of the window come
    back as 0 like ``pad_input`` leaves them.  Returns ``(action_log_probs [B, A], entropy [B, A] or None)``.
    """
    if hidden_packed.dim() == 3:
        if hidden_packed.shape[0] != 1:
            raise ValueError(f"packed hidden must be [nnz, H] or [1, nnz, H], got {tuple(hidden_packed.shape)}")
        hidden_packed = hidden_packed[0]
    if sequences.dim() != 2 or attention_mask.shape != sequences.shape:
        raise ValueError("sequences and attention_mask must both be [B, S]")
    B, S = sequences.shape
    A = int(num_actions)
    if not 0 < A < S:
        raise ValueError(f"num_actions must be in (0, {S}), got {A}")
    dev = hidden_packed.device
    valid = attention_mask.to(dev).reshape(-1) != 0
    nnz_positions = torch.nonzero(valid).squeeze(-1)          # padded flat position of every packed token
    if hidden_packed.shape[0] != nnz_positions.numel():
        raise ValueError(f"hidden_packed has {hidden_packed.shape[0]} rows but attention_mask has "
                         f"{nnz_positions.numel()} valid tokens")
    packed_tokens = sequences.to(dev).reshape(-1)[nnz_positions]
    rolled = torch.roll(packed_tokens, shifts=-1, dims=0)
    packed_index = torch.cumsum(valid.long(), 0) - 1          # packed row of every valid padded position
    window = torch.zeros((B, S), dtype=torch.bool, device=dev)
    window[:, S - A - 1:S - 1] = True
    picked = torch.nonzero(window.reshape(-1) & valid).squeeze(-1)
    rows = packed_index[picked]
    lp, ent = fused_lm_head(hidden_packed[rows], lm_head_weight, rolled[rows], temperature=temperature,
                            compute_entropy=compute_entropy, chunk=chunk)

    def unpack(x):
        full = x.new_zeros(B * S).index_copy(0, picked, x)
        return full.view(B, S)[:, S - A - 1:S - 1]

    return unpack(lp), (unpack(ent) if compute_entropy else None)
