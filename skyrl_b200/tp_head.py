"""Vocab-parallel (tensor-parallel) fused LM head: the vocabulary is sharded over the GPUs of a TP group.

B200 counterpart of the reference's Megatron helpers ``DistributedLogprob`` / ``_VocabParallelEntropy``
(skyrl-train/skyrl_train/distributed/megatron/model_utils.py:64-247, 548-590).  The reference all-reduces three
``[N, 1]`` tensors for log-probs (max, sum-exp, gathered label logit) and three more for entropy
(model_utils.py:44-59, 555-563); here every rank runs the fused head on its own ``[V/tp, H]`` shard (logits never
materialised), the ranks exchange **one** ``[N, 4]`` fp32 tensor (all-gather over NVLink) and combine locally.
Backward: each rank forms d(logits) for its shard from the *global* log-sum-exp, produces its ``d_weight`` shard
locally, and the partial ``d_hidden`` are summed with one all-reduce.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib

def combine_shard_stats(stats: torch.Tensor, temperature: float) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """``stats`` [tp, N, 4] = per shard {max, sum, weighted sum, label logit} in log2 units (CUDA tensor).
    Returns ``(logprobs, entropy, lse2)`` [N] fp32 — the same numbers the unsharded head produces."""
    if not stats.is_cuda:
        raise RuntimeError("combine_shard_stats needs CUDA tensors (no CPU fallback)")
    lib = _lib.load()
    stats = stats.contiguous()
    tp, N = int(stats.shape[0]), int(stats.shape[1])
    with torch.cuda.device(stats.device):
        lp = torch.empty(N, dtype=torch.float32, device=stats.device)
        ent = torch.empty_like(lp)
        lse2 = torch.empty_like(lp)
        _lib.check(lib.skyb200_head_combine_stats(_lib.ptr(stats), tp, N, float(temperature), _lib.ptr(lp),
                                                  _lib.ptr(ent), _lib.ptr(lse2), _lib.stream_ptr()),
                   "skyb200_head_combine_stats")
    return lp, ent, lse2


def shard_forward_stats(hidden: torch.Tensor, weight_shard: torch.Tensor, labels_local: torch.Tensor,
                        temperature: float, chunk: int = 0) -> torch.Tensor:
    """This rank's ``[N, 4]`` stats (one C-ABI call: tcgen05 GEMM + online-softmax epilogue over the shard)."""
    lib = _lib.load()
    dev = hidden.device
    N, H = hidden.shape
    V = weight_shard.shape[0]
    with torch.cuda.device(dev):
        stats = torch.empty((N, 4), dtype=torch.float32, device=dev)
        wsb = lib.skyb200_head_forward_workspace(N, H, V, chunk)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        _lib.check(lib.skyb200_head_forward_stats(_lib.ptr(hidden), _lib.ptr(weight_shard), _lib.ptr(labels_local),
                                                  _lib.ptr(stats), _lib.ptr(ws), wsb, N, H, V, float(temperature),
                                                  int(chunk), _lib.stream_ptr()), "skyb200_head_forward_stats")
    return stats


def shard_backward(hidden, weight_shard, labels_local, lse2, entropy, g_logprobs, g_entropy, temperature,
                   need_d_hidden=True, need_d_weight=True, chunk: int = 0):
    """This rank's partial ``d_hidden`` [N, H] bf16 and its ``d_weight`` shard [V/tp, H] fp32."""
    lib = _lib.load()
    dev = hidden.device
    N, H = hidden.shape
    V = weight_shard.shape[0]
    with torch.cuda.device(dev):
        d_hidden = torch.empty((N, H), dtype=torch.bfloat16, device=dev) if need_d_hidden else None
        d_weight = torch.empty((V, H), dtype=torch.float32, device=dev) if need_d_weight else None
        wsb = lib.skyb200_head_backward_workspace(N, H, V, chunk)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        _lib.check(
            lib.skyb200_head_backward(_lib.ptr(hidden), _lib.ptr(weight_shard), _lib.ptr(labels_local), _lib.ptr(lse2),
                                      _lib.ptr(entropy) if g_entropy is not None else None, _lib.ptr(g_logprobs),
                                      _lib.ptr(g_entropy), _lib.ptr(d_hidden), _lib.ptr(d_weight), 0, _lib.ptr(ws),
                                      wsb, N, H, V, float(temperature), int(chunk), _lib.stream_ptr()),
            "skyb200_head_backward",
        )
    return d_hidden, d_weight


class VocabParallelHeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hidden, weight_shard, labels, vocab_start, temperature, group, compute_entropy):
        if not hidden.is_cuda:
            raise RuntimeError("vocab_parallel_lm_head needs CUDA tensors (no CPU fallback)")
        h = hidden.detach()
        h = (h if h.dtype == torch.bfloat16 else h.to(torch.bfloat16)).contiguous()
        w = weight_shard.detach()
        w = (w if w.dtype == torch.bfloat16 else w.to(torch.bfloat16)).contiguous()
        lab_local = (labels.detach().reshape(-1).to(device=h.device, dtype=torch.int64) - int(vocab_start)).contiguous()
        stats = shard_forward_stats(h, w, lab_local, temperature)
        tp = dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1
        if tp > 1:
            gathered = torch.empty((tp,) + tuple(stats.shape), dtype=stats.dtype, device=stats.device)
            dist.all_gather_into_tensor(gathered, stats, group=group)  # 16 bytes per token per rank
        else:
            gathered = stats.unsqueeze(0)
        logprobs, entropy, lse2 = combine_shard_stats(gathered, temperature)
        ctx.save_for_backward(h, w, lab_local, lse2.contiguous(), entropy.contiguous())
        ctx.meta = (float(temperature), group, tp, hidden.dtype, weight_shard.dtype, bool(compute_entropy))
        return logprobs, (entropy if compute_entropy else torch.empty(0, device=h.device))

    @staticmethod
    def backward(ctx, g_logprobs, g_entropy):
        h, w, lab_local, lse2, entropy = ctx.saved_tensors
        T, group, tp, h_dtype, w_dtype, compute_entropy = ctx.meta
        N = h.shape[0]
        g_lp = (g_logprobs.detach().to(torch.float32).contiguous() if g_logprobs is not None
                else torch.zeros(N, dtype=torch.float32, device=h.device))
        g_ent = None
        if compute_entropy and g_entropy is not None and g_entropy.numel() == N:
            g_ent = g_entropy.detach().to(torch.float32).contiguous()
        need_dh, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        d_hidden, d_weight = shard_backward(h, w, lab_local, lse2, entropy, g_lp, g_ent, T, need_dh, need_dw)
        if need_dh and tp > 1:
            acc = d_hidden.float()  # sum the tp partial products in fp32
            dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
            d_hidden = acc.to(torch.bfloat16)
        gh = None if not need_dh else (d_hidden if h_dtype == torch.bfloat16 else d_hidden.to(h_dtype))
        gw = None if not need_dw else (d_weight if w_dtype == torch.float32 else d_weight.to(w_dtype))
        return gh, gw, None, None, None, None, None


def vocab_parallel_lm_head(hidden: torch.Tensor, weight_shard: torch.Tensor, labels: torch.Tensor, vocab_start: int,
                           temperature: float = 1.0, group=None, compute_entropy: bool = True):
    """log p(label) and entropy per row with the vocabulary sharded over ``group``.

    ``hidden`` [N, H] (replicated across the TP group), ``weight_shard`` [V/tp, H] = rows
    ``[vocab_start, vocab_start + V/tp)`` of the lm_head weight, ``labels`` [N] GLOBAL token ids.
    """
    lp, ent = VocabParallelHeadFn.apply(hidden, weight_shard, labels, vocab_start, temperature, group, compute_entropy)
    return lp, (ent if compute_entropy else None)
