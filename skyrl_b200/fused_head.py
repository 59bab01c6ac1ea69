"""Fused LM head: projection + online log-softmax + label gather + entropy, forward and backward.

Replaces, for the policy update, the chain the reference runs at model_wrapper.py:309-363:
HF ``lm_head`` GEMM -> ``logits.div_(temperature)`` -> ``logprobs_from_logits`` (flash-attn Triton cross-entropy)
-> ``chunked_entropy_from_logits``.  The [tokens x vocab] logits never exist: a tcgen05 GEMM reduces every
128x256 logits tile to online-softmax partials while it is still in tensor memory.

    logprobs, entropy = fused_lm_head(hidden, weight, labels, temperature=1.0, entropy_mask=None)

``hidden`` [N, H] bf16 (rows = the response positions you need; pack them, do not pad), ``weight`` [V, H]
(``lm_head.weight``, bf16; an fp32 master weight is cast once), ``labels`` [N] int64.
Outputs are fp32 [N].  Gradients: ``d_hidden`` bf16 and ``d_weight`` accumulated in fp32.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib


def _workspace(nbytes: int, dev: torch.device) -> torch.Tensor:
    return torch.empty(nbytes, dtype=torch.uint8, device=dev)


class FusedLMHeadFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hidden, weight, labels, temperature, entropy_mask, compute_entropy, chunk):
        lib = _lib.load()
        if not hidden.is_cuda:
            raise RuntimeError("fused_lm_head needs CUDA tensors (no CPU fallback)")
        if hidden.dim() != 2 or weight.dim() != 2 or hidden.shape[1] != weight.shape[1]:
            raise ValueError(f"expected hidden [N, H] and weight [V, H], got {tuple(hidden.shape)} {tuple(weight.shape)}")
        dev = hidden.device
        N, H = hidden.shape
        V = weight.shape[0]
        h = hidden.detach()
        if h.dtype != torch.bfloat16:
            h = h.to(torch.bfloat16)
        h = h.contiguous()
        w = weight.detach()
        if w.dtype != torch.bfloat16:
            w = w.to(torch.bfloat16)
        w = w.contiguous()
        lab = labels.detach().reshape(-1).to(device=dev, dtype=torch.int64).contiguous()
        if lab.numel() != N:
            raise ValueError(f"labels has {lab.numel()} entries for {N} hidden rows")
        emask = None
        if entropy_mask is not None:
            emask = entropy_mask.detach().reshape(-1).to(device=dev, dtype=torch.float32).contiguous()
        with torch.cuda.device(dev):
            logprobs = torch.empty(N, dtype=torch.float32, device=dev)
            entropy = torch.empty(N, dtype=torch.float32, device=dev) if compute_entropy else None
            lse2 = torch.empty(N, dtype=torch.float32, device=dev)
            ws_bytes = lib.skyb200_head_forward_workspace(N, H, V, chunk)
            ws = _workspace(ws_bytes, dev)
            # the unmasked entropy is what backward needs; the mask is applied on the way out
            _lib.check(
                lib.skyb200_head_forward(_lib.ptr(h), _lib.ptr(w), _lib.ptr(lab), None, _lib.ptr(logprobs),
                                         _lib.ptr(entropy), _lib.ptr(lse2), _lib.ptr(ws), ws_bytes, N, H, V,
                                         float(temperature), int(chunk), _lib.stream_ptr()),
                "skyb200_head_forward",
            )
        ctx.temperature = float(temperature)
        ctx.chunk = int(chunk)
        ctx.compute_entropy = bool(compute_entropy)
        ctx.hidden_dtype = hidden.dtype
        ctx.weight_ref = weight
        ctx.weight_dtype = weight.dtype
        empty = torch.empty(0, device=dev)
        ctx.save_for_backward(h, w, lab, lse2, entropy if compute_entropy else empty,
                              emask if emask is not None else empty)
        ctx.has_mask = emask is not None
        if compute_entropy:
            ent_out = entropy * emask if emask is not None else entropy
        else:
            ent_out = empty
        return logprobs, ent_out

    @staticmethod
    def backward(ctx, g_logprobs, g_entropy):
        lib = _lib.load()
        h, w, lab, lse2, entropy, emask = ctx.saved_tensors
        dev = h.device
        N, H = h.shape
        V = w.shape[0]
        need_dh, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (need_dh or need_dw):
            return None, None, None, None, None, None, None
        if g_logprobs is None:
            g_lp = torch.zeros(N, dtype=torch.float32, device=dev)
        else:
            g_lp = g_logprobs.detach().to(torch.float32).contiguous()
        g_ent = None
        if ctx.compute_entropy and g_entropy is not None and g_entropy.numel() == N:
            g_ent = g_entropy.detach().to(torch.float32)
            if ctx.has_mask:
                g_ent = g_ent * emask
            g_ent = g_ent.contiguous()
        main_grad = getattr(ctx.weight_ref, "main_grad", None)
        with torch.cuda.device(dev):
            d_hidden = torch.empty((N, H), dtype=torch.bfloat16, device=dev) if need_dh else None
            accumulate = 0
            if need_dw:
                if main_grad is not None and main_grad.dtype == torch.float32 and main_grad.shape == (V, H):
                    d_weight, accumulate = main_grad, 1  # Megatron-style fp32 main_grad: accumulate in place
                else:
                    d_weight = torch.empty((V, H), dtype=torch.float32, device=dev)
            else:
                d_weight = None
            ws_bytes = lib.skyb200_head_backward_workspace(N, H, V, ctx.chunk)
            ws = _workspace(ws_bytes, dev)
            _lib.check(
                lib.skyb200_head_backward(_lib.ptr(h), _lib.ptr(w), _lib.ptr(lab), _lib.ptr(lse2),
                                          _lib.ptr(entropy) if g_ent is not None else None, _lib.ptr(g_lp),
                                          _lib.ptr(g_ent), _lib.ptr(d_hidden), _lib.ptr(d_weight), accumulate,
                                          _lib.ptr(ws), ws_bytes, N, H, V, ctx.temperature, ctx.chunk,
                                          _lib.stream_ptr()),
                "skyb200_head_backward",
            )
        gh = None
        if need_dh:
            gh = d_hidden if ctx.hidden_dtype == torch.bfloat16 else d_hidden.to(ctx.hidden_dtype)
        gw = None
        if need_dw and accumulate == 0:
            gw = d_weight if ctx.weight_dtype == torch.float32 else d_weight.to(ctx.weight_dtype)
        return gh, gw, None, None, None, None, None


def fused_lm_head(
    hidden: torch.Tensor,
    weight: torch.Tensor,
    labels: torch.Tensor,
    temperature: float = 1.0,
    entropy_mask: Optional[torch.Tensor] = None,
    compute_entropy: bool = True,
    chunk: int = 0,
) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """log p(label) and entropy per row of ``hidden`` without materialising logits. See module docstring."""
    logprobs, entropy = FusedLMHeadFunction.apply(hidden, weight, labels, temperature, entropy_mask,
                                                  compute_entropy, chunk)
    return logprobs, (entropy if compute_entropy else None)


@torch.no_grad()
def fused_lm_head_logprobs(hidden, weight, labels, temperature: float = 1.0, chunk: int = 0) -> torch.Tensor:
    """No-grad forward (the old-policy / reference-policy logprob pass, worker.py:979-1000)."""
    lp, _ = FusedLMHeadFunction.apply(hidden, weight, labels, temperature, None, False, chunk)
    return lp
