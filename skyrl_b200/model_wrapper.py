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


class _SeqParallelGather(torch.autograd.Function):
    """All-gather of ``[B, L]`` shards along the sequence dimension.  Mirrors the reference's ``Gather``
    (skyrl_train/distributed/ulysses/utils.py:199-239): backward hands every rank its own slice of the upstream
    gradient, multiplied by the group size when ``grad_scaler`` (the reference's default)."""

    @staticmethod
    def forward(ctx, local, group, grad_scaler):
        import torch.distributed as dist

        ctx.ws, ctx.rank = dist.get_world_size(group), dist.get_rank(group)
        ctx.scale, ctx.part = bool(grad_scaler), local.shape[1]
        parts = [torch.empty_like(local) for _ in range(ctx.ws)]
        dist.all_gather(parts, local.contiguous(), group=group)
        return torch.cat(parts, dim=1)

    @staticmethod
    def backward(ctx, g):
        if ctx.scale:
            g = g * ctx.ws
        return g.split(ctx.part, dim=1)[ctx.rank].contiguous(), None, None


class _ZeroAnchor(torch.autograd.Function):
    """``zeros([B, L])`` that still depends on ``hidden_local``: a rank whose slice holds no response position must
    run the trunk's backward too (its attention backward is a collective), with a zero gradient - which is what the
    reference's sliced-away log-probs deliver."""

    @staticmethod
    def forward(ctx, hidden_local):
        ctx.meta = (hidden_local.shape, hidden_local.dtype)
        return hidden_local.new_zeros(hidden_local.shape[:2], dtype=torch.float32)

    @staticmethod
    def backward(ctx, g):
        shape, dtype = ctx.meta
        return g.new_zeros(shape, dtype=dtype)


def action_log_probs_from_sp_hidden(
    hidden_local: torch.Tensor,
    lm_head_weight: torch.Tensor,
    sequences: torch.Tensor,
    num_actions: int,
    attention_mask: Optional[torch.Tensor] = None,
    temperature: float = 1.0,
    compute_entropy: bool = False,
    group=None,
    grad_scaler: bool = True,
    chunk: int = 0,
) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """The Ulysses sequence-parallel branch of the wrapper (model_wrapper.py:292-305, 319-355; padded layout).

    The trunk ran on this rank's slice of the sequence dimension (``ulysses_pad_and_slice_inputs``: ``S`` padded to a
    multiple of the group size, rank r owns columns ``[r*L, (r+1)*L)``); ``hidden_local`` [B, L, H] is its last hidden
    state, ``sequences`` / ``attention_mask`` are the full [B, S] tensors every rank holds.  The head is per token,
    so each rank runs the fused head on the response-window rows **of its own slice only** and the ranks all-gather
    the fp32 log-probs (4 bytes per position; the reference gathers the same tensor after projecting all ``S/sp``
    positions to the vocabulary).  Returns ``(action_log_probs [B, A], entropy [B, A] or None)`` on every rank.
    """
    import torch.distributed as dist

    if hidden_local.dim() != 3 or sequences.dim() != 2 or hidden_local.shape[0] != sequences.shape[0]:
        raise ValueError(f"expected hidden_local [B, S/sp, H] and sequences [B, S], got {tuple(hidden_local.shape)} "
                         f"{tuple(sequences.shape)}")
    sp, rank = dist.get_world_size(group), dist.get_rank(group)
    B, L, H = hidden_local.shape
    S = sequences.shape[1]
    A = int(num_actions)
    if not 0 < A < S:
        raise ValueError(f"num_actions must be in (0, {S}), got {A}")
    pad = (sp - S % sp) % sp
    if L * sp != S + pad:
        raise ValueError(f"hidden_local has {L} positions per rank, expected {(S + pad) // sp} (= ceil({S} / {sp}))")
    if compute_entropy and attention_mask is not None and attention_mask.shape != (B, S):
        raise ValueError(
            f"attention_mask shape {attention_mask.shape} does not match logits shape "
            f"(batch_size={B}, seqlen={S}). Expected attention_mask shape: ({B}, {S})"
        )
    lo, hi = S - A - 1, S - 1                                  # the response window in global columns
    g0 = rank * L
    c0, c1 = min(max(lo - g0, 0), L), min(max(hi - g0, 0), L)  # ... and in this rank's columns
    n = c1 - c0
    lp_local = ent_local = _ZeroAnchor.apply(hidden_local)
    if n > 0:
        rolled = torch.roll(sequences, shifts=-1, dims=1)[:, g0 + c0:g0 + c1]
        emask = None
        if compute_entropy and attention_mask is not None:
            emask = attention_mask[:, g0 + c0:g0 + c1].reshape(B * n)
        lp, ent = fused_lm_head(hidden_local[:, c0:c1, :].reshape(B * n, H), lm_head_weight, rolled.reshape(B * n),
                                temperature=temperature, entropy_mask=emask, compute_entropy=compute_entropy,
                                chunk=chunk)
        place = lambda x: torch.nn.functional.pad(x.view(B, n), (c0, L - c1))
        lp_local = lp_local + place(lp)
        if compute_entropy:
            ent_local = ent_local + place(ent)
    full = _SeqParallelGather.apply(lp_local, group, grad_scaler)
    out_ent = None
    if compute_entropy:
        out_ent = _SeqParallelGather.apply(ent_local, group, grad_scaler)[:, lo:hi]
    return full[:, lo:hi], out_ent


def action_log_probs_from_sp_packed_hidden(
    hidden_local: torch.Tensor,
    lm_head_weight: torch.Tensor,
    sequences: torch.Tensor,
    attention_mask: torch.Tensor,
    num_actions: int,
    temperature: float = 1.0,
    compute_entropy: bool = False,
    group=None,
    grad_scaler: bool = True,
    chunk: int = 0,
) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Sample packing **and** Ulysses sequence parallelism (model_wrapper.py:278-305, 319-358): the un-padded token
    stream ``[1, nnz]`` was padded to a multiple of the group size and rank r ran the trunk on packed positions
    ``[r*L, (r+1)*L)``; ``hidden_local`` is ``[L, H]`` (or ``[1, L, H]``).

    Each rank projects only the packed rows of its slice that land in a response window ``[-A-1:-1]`` of their
    sequence (labels: the packed stream rolled by one *before* slicing, as the reference does), the ranks all-gather
    the fp32 results along the packed dimension, and every rank pads them back to ``[B, S]`` (zeros at masked
    positions, like ``pad_input``) and slices.  Returns ``(action_log_probs [B, A], entropy [B, A] or None)``.
    """
    import torch.distributed as dist

    if hidden_local.dim() == 3:
        if hidden_local.shape[0] != 1:
            raise ValueError(f"packed hidden must be [L, H] or [1, L, H], got {tuple(hidden_local.shape)}")
        hidden_local = hidden_local[0]
    if sequences.dim() != 2 or attention_mask.shape != sequences.shape:
        raise ValueError("sequences and attention_mask must both be [B, S]")
    sp, rank = dist.get_world_size(group), dist.get_rank(group)
    B, S = sequences.shape
    A = int(num_actions)
    if not 0 < A < S:
        raise ValueError(f"num_actions must be in (0, {S}), got {A}")
    dev = hidden_local.device
    L = hidden_local.shape[0]
    valid = attention_mask.to(dev).reshape(-1) != 0
    nnz_positions = torch.nonzero(valid).squeeze(-1)          # padded flat position of every packed token
    nnz = int(nnz_positions.numel())
    pad = (sp - nnz % sp) % sp
    if L * sp != nnz + pad:
        raise ValueError(f"hidden_local has {L} rows per rank, expected {(nnz + pad) // sp} (= ceil({nnz} / {sp}))")
    packed_tokens = sequences.to(dev).reshape(-1)[nnz_positions]
    rolled = torch.roll(packed_tokens, shifts=-1, dims=0)     # rolled over the whole stream, then sliced
    window = torch.zeros((B, S), dtype=torch.bool, device=dev)
    window[:, S - A - 1:S - 1] = True
    in_window = window.reshape(-1)[nnz_positions]             # per packed token
    g0 = rank * L
    mine = torch.nonzero(in_window[g0:min(g0 + L, nnz)]).squeeze(-1)   # local rows that reach the head
    anchor = _ZeroAnchor.apply(hidden_local.unsqueeze(0))              # [1, L], keeps the trunk's backward alive
    lp_local = ent_local = anchor
    if mine.numel() > 0:
        lp, ent = fused_lm_head(hidden_local[mine], lm_head_weight, rolled[g0 + mine], temperature=temperature,
                                compute_entropy=compute_entropy, chunk=chunk)
        place = lambda x: x.new_zeros(L).index_copy(0, mine, x).unsqueeze(0)
        lp_local = anchor + place(lp)
        if compute_entropy:
            ent_local = anchor + place(ent)

    def finish(local):
        stream = _SeqParallelGather.apply(local, group, grad_scaler)[0, :nnz]      # [nnz]
        full = stream.new_zeros(B * S).index_copy(0, nnz_positions, stream)       # pad_input
        return full.view(B, S)[:, S - A - 1:S - 1]

    return finish(lp_local), (finish(ent_local) if compute_entropy else None)
