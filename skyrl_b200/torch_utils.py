"""CUDA versions of the reference's logits helpers, same names and call signatures.

Reference: skyrl-train/skyrl_train/utils/torch_utils.py —
``logprobs_from_logits`` (:115-145), ``logprobs_from_logits_v2`` (:158-177),
``chunked_entropy_from_logits`` (:59-111), ``masked_mean`` (:180-184), ``safe_exp_delta`` (:187-192).

These operate on *materialised* logits (for callers that still run the HF lm_head).  The fused path that never
builds the logits is :mod:`skyrl_b200.fused_head`.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib

_DTYPE_CODE = {torch.float32: 0, torch.bfloat16: 1}

# Storages whose logits are still needed by the backward of a grad-requiring entropy node (storage pointer -> count).
# ``HFModelWrapper.forward`` calls ``logprobs_from_logits(logits, inplace_backward=True)`` AND
# ``chunked_entropy_from_logits(logits, requires_grad=use_entropy_loss)`` on the SAME buffer (model_wrapper.py:317,
# 347): if the log-prob node's backward overwrote that buffer with d_logits before the entropy node's backward has
# read it, the entropy gradient would be silently wrong.  (The reference is safe because its entropy goes through
# ``log_softmax``, which saves its OUTPUT.)  The log-prob backward therefore falls back to a fresh d_logits buffer while
# an entropy node on the same storage is pending, and every in-place overwrite bumps the tensor's version counter so
# that any OTHER consumer that saved the logits gets autograd's "modified by an inplace operation" error, not garbage.
_PENDING_ENTROPY: dict = {}


def _storage_key(t: torch.Tensor) -> int:
    return t.untyped_storage().data_ptr()


def _pending_add(key: int) -> None:
    _PENDING_ENTROPY[key] = _PENDING_ENTROPY.get(key, 0) + 1


def _pending_drop(key: int) -> None:
    n = _PENDING_ENTROPY.get(key, 0) - 1
    if n > 0:
        _PENDING_ENTROPY[key] = n
    else:
        _PENDING_ENTROPY.pop(key, None)


def _check_logits(logits: torch.Tensor) -> torch.Tensor:
    if not logits.is_cuda:
        raise RuntimeError("skyrl_b200.torch_utils needs CUDA logits (no CPU fallback)")
    if logits.dtype not in _DTYPE_CODE:
        logits = logits.float()
    return logits


class _LogitsLogprobEntropy(torch.autograd.Function):
    """(logprobs, entropy) from logits [N, V] in one read; backward writes d_logits (optionally in place)."""

    @staticmethod
    def forward(ctx, logits2d, labels1d, entropy_mask1d, want_logprobs, want_entropy, inplace_backward):
        lib = _lib.load()
        N, V = logits2d.shape
        dev = logits2d.device
        with torch.cuda.device(dev):
            lp = torch.empty(N, dtype=torch.float32, device=dev) if want_logprobs else None
            ent = torch.empty(N, dtype=torch.float32, device=dev) if want_entropy else None
            lse = torch.empty(N, dtype=torch.float32, device=dev)
            # the kernel applies the mask; keep the unmasked entropy for backward only if a grad is needed
            _lib.check(
                lib.skyb200_logits_logprob_entropy(_lib.ptr(logits2d), _DTYPE_CODE[logits2d.dtype], _lib.ptr(labels1d),
                                                   None, _lib.ptr(lp), _lib.ptr(ent), _lib.ptr(lse), N, V,
                                                   _lib.stream_ptr()),
                "skyb200_logits_logprob_entropy",
            )
        ctx.inplace = bool(inplace_backward)
        ctx.has_mask = entropy_mask1d is not None
        ctx.pending_key = None
        if want_entropy and ctx.needs_input_grad[0]:
            import weakref

            ctx.pending_key = _storage_key(logits2d)
            _pending_add(ctx.pending_key)
            # a graph that is dropped without running its backward must not leave the entry behind
            ctx.pending_fin = weakref.finalize(ctx, _pending_drop, ctx.pending_key)
        saved_ent = ent if want_entropy else torch.empty(0, device=dev)
        ctx.save_for_backward(logits2d, labels1d if labels1d is not None else torch.empty(0, device=dev), lse,
                              saved_ent, entropy_mask1d if entropy_mask1d is not None else torch.empty(0, device=dev))
        ctx.want = (want_logprobs, want_entropy)
        ent_out = ent
        if want_entropy and entropy_mask1d is not None:
            ent_out = ent * entropy_mask1d
        outs = (lp if want_logprobs else torch.empty(0, device=dev),
                ent_out if want_entropy else torch.empty(0, device=dev))
        return outs

    @staticmethod
    def backward(ctx, g_lp, g_ent):
        lib = _lib.load()
        logits2d, labels1d, lse, ent, emask = ctx.saved_tensors
        want_lp, want_ent = ctx.want
        N, V = logits2d.shape
        dev = logits2d.device
        g_lp_c = g_lp.float().contiguous() if want_lp and g_lp is not None else None
        g_ent_c = None
        if want_ent and g_ent is not None:
            g_ent_c = g_ent.float()
            if ctx.has_mask:
                g_ent_c = g_ent_c * emask
            g_ent_c = g_ent_c.contiguous()
        if ctx.pending_key is not None:  # this entropy node has now read the logits
            ctx.pending_fin.detach()
            _pending_drop(ctx.pending_key)
        inplace = ctx.inplace and _PENDING_ENTROPY.get(_storage_key(logits2d), 0) == 0
        with torch.cuda.device(dev):
            d_logits = logits2d if inplace else torch.empty_like(logits2d)
            _lib.check(
                lib.skyb200_logits_backward(_lib.ptr(logits2d), _DTYPE_CODE[logits2d.dtype],
                                            _lib.ptr(labels1d) if labels1d.numel() else None, _lib.ptr(lse),
                                            _lib.ptr(ent) if want_ent else None, _lib.ptr(g_lp_c), _lib.ptr(g_ent_c),
                                            _lib.ptr(d_logits), N, V, _lib.stream_ptr()),
                "skyb200_logits_backward",
            )
        if inplace:  # the raw-pointer write is invisible to autograd: make it visible
            torch.autograd.graph.increment_version(logits2d)
        return d_logits, None, None, None, None, None


def logprobs_from_logits(logits: torch.Tensor, labels: torch.Tensor, inplace_backward: bool = True) -> torch.Tensor:
    """Per-token log-probability of ``labels`` (reference: torch_utils.py:115-145).

    ``logits`` [..., V] bf16/fp32 on CUDA, ``labels`` [...] int64.  Result fp32 of shape ``logits.shape[:-1]``.
    With ``inplace_backward=True`` (the reference default) backward overwrites the logits buffer with d_logits.
    """
    logits = _check_logits(logits)
    batch_dim = logits.shape[:-1]
    l2 = logits.reshape(-1, logits.shape[-1])
    if not l2.is_contiguous():
        l2 = l2.contiguous()
    lab = labels.reshape(-1).to(device=l2.device, dtype=torch.int64).contiguous()
    lp, _ = _LogitsLogprobEntropy.apply(l2, lab, None, True, False, inplace_backward)
    return lp.view(*batch_dim)


def logprobs_from_logits_v2(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """Same contract as the reference's memory-efficient variant (torch_utils.py:158-177)."""
    return logprobs_from_logits(logits, labels, inplace_backward=False)


def chunked_entropy_from_logits(
    logits: torch.Tensor,
    requires_grad: bool = False,
    attention_mask: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """Per-position softmax entropy (reference: torch_utils.py:59-111).

    ``logits`` [B, S, V]; ``attention_mask`` [B, S] zeroes padded positions; raises ``ValueError`` on a mask
    shape mismatch like the reference.  The result has the logits' dtype (the reference stores into
    ``zeros(..., dtype=logits.dtype)``); the arithmetic is fp32.
    """
    if attention_mask is not None and attention_mask.shape != (logits.shape[0], logits.shape[1]):
        raise ValueError(
            f"attention_mask shape {attention_mask.shape} does not match logits shape "
            f"(batch_size={logits.shape[0]}, seqlen={logits.shape[1]}). "
            f"Expected attention_mask shape: ({logits.shape[0]}, {logits.shape[1]})"
        )
    out_dtype = logits.dtype
    logits = _check_logits(logits)
    B, S, V = logits.shape
    l2 = logits.reshape(-1, V)
    if not l2.is_contiguous():
        l2 = l2.contiguous()
    mask = None
    if attention_mask is not None:
        mask = attention_mask.reshape(-1).to(device=l2.device, dtype=torch.float32).contiguous()
    if requires_grad:
        _, ent = _LogitsLogprobEntropy.apply(l2, None, mask, False, True, False)
    else:
        with torch.no_grad():
            _, ent = _LogitsLogprobEntropy.apply(l2, None, mask, False, True, False)
    return ent.view(B, S).to(out_dtype)


def masked_mean(tensor: torch.Tensor, mask: Optional[torch.Tensor], dim: Optional[int] = None) -> torch.Tensor:
    """``(tensor*mask).sum(dim) / mask.sum(dim).clamp(min=1)`` (reference: torch_utils.py:180-184).

    The full reduction (``dim=None``) of a no-grad CUDA fp32 tensor runs in libskyb200; everything else
    (``dim`` given, autograd needed) is expressed with torch ops so gradients flow as in the reference.
    """
    if dim is None and tensor.is_cuda and not tensor.requires_grad and tensor.dtype == torch.float32:
        lib = _lib.load()
        x = tensor.contiguous()
        m = None if mask is None else mask.to(device=x.device, dtype=torch.float32).contiguous()
        with torch.cuda.device(x.device):
            out = torch.empty((), dtype=torch.float32, device=x.device)
            ws = torch.empty(16384, dtype=torch.uint8, device=x.device)
            _lib.check(lib.skyb200_masked_mean(_lib.ptr(x), _lib.ptr(m), _lib.ptr(out), _lib.ptr(ws), ws.numel(),
                                               x.numel(), _lib.stream_ptr()), "skyb200_masked_mean")
        return out
    if mask is None:
        return tensor.mean(axis=dim)
    return (tensor * mask).sum(axis=dim) / mask.sum(axis=dim).clamp(min=1.0)


def safe_exp_delta(delta: torch.Tensor, clip: float = 20.0, out_dtype=None) -> torch.Tensor:
    """``exp(clamp(delta, -clip, clip))`` in fp32 (reference: torch_utils.py:187-192)."""
    y = torch.clamp(delta.to(torch.float32), -clip, clip).exp()
    return y.to(out_dtype or delta.dtype)
