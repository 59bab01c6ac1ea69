"""Critic value head on the GPU: ``values = hidden @ w + b`` with autograd.

Reference: the critic wrapper applies ``nn.Linear(hidden_size, 1)`` to the trunk's last hidden state and keeps
``values[:, :-1][:, -num_actions:]`` (skyrl-train/skyrl_train/model_wrapper.py:472-489).  :func:`action_values`
does the slicing first, so only the A response rows per sequence are projected.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib


class _ValueHeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hidden2d, weight, bias):
        lib = _lib.load()
        if not hidden2d.is_cuda:
            raise RuntimeError("skyrl_b200 value head needs CUDA tensors (no CPU fallback)")
        dev = hidden2d.device
        N, H = hidden2d.shape
        h = hidden2d.detach()
        h = (h if h.dtype == torch.bfloat16 else h.to(torch.bfloat16)).contiguous()
        w = weight.detach().reshape(-1).to(device=dev, dtype=torch.float32).contiguous()
        if w.numel() != H:
            raise ValueError(f"value head weight has {w.numel()} entries for hidden size {H}")
        b = float(bias.detach().reshape(-1)[0].item()) if bias is not None else 0.0
        with torch.cuda.device(dev):
            values = torch.empty(N, dtype=torch.float32, device=dev)
            _lib.check(lib.skyb200_value_head_forward(_lib.ptr(h), _lib.ptr(w), b, _lib.ptr(values), N, H,
                                                      _lib.stream_ptr()), "skyb200_value_head_forward")
        ctx.save_for_backward(h, w)
        ctx.meta = (hidden2d.dtype, weight.dtype, tuple(weight.shape), bias is not None)
        return values

    @staticmethod
    def backward(ctx, g_values):
        lib = _lib.load()
        h, w = ctx.saved_tensors
        h_dtype, w_dtype, w_shape, has_bias = ctx.meta
        dev = h.device
        N, H = h.shape
        g = g_values.detach().to(torch.float32).contiguous()
        need_dh, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        with torch.cuda.device(dev):
            d_hidden = torch.empty((N, H), dtype=torch.bfloat16, device=dev) if need_dh else None
            d_w = torch.empty(H, dtype=torch.float32, device=dev) if need_dw else None
            wsb = lib.skyb200_value_head_backward_workspace(N, H)
            ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
            _lib.check(lib.skyb200_value_head_backward(_lib.ptr(h), _lib.ptr(w), _lib.ptr(g), _lib.ptr(d_hidden),
                                                       _lib.ptr(d_w), 0, _lib.ptr(ws), wsb, N, H, _lib.stream_ptr()),
                       "skyb200_value_head_backward")
        gh = None if d_hidden is None else (d_hidden if h_dtype == torch.bfloat16 else d_hidden.to(h_dtype))
        gw = None if d_w is None else d_w.to(w_dtype).view(w_shape)
        gb = g.sum().reshape(1) if (has_bias and ctx.needs_input_grad[2]) else None
        return gh, gw, gb


def value_head(hidden: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``hidden`` [..., H] bf16, ``weight`` [1, H] or [H], ``bias`` [1] or None -> fp32 values of shape ``[...]``."""
    lead = hidden.shape[:-1]
    return _ValueHeadFn.apply(hidden.reshape(-1, hidden.shape[-1]), weight, bias).view(*lead)


def action_values(hidden: torch.Tensor, weight: torch.Tensor, num_actions: int,
                  bias: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``values[:, :-1][:, -num_actions:]`` of the reference (model_wrapper.py:489-495) from ``hidden`` [B, S, H]."""
    S = hidden.shape[1]
    A = int(num_actions)
    return value_head(hidden[:, S - 1 - A:S - 1, :], weight, bias)
