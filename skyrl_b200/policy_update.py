"""One policy-update step of the hot path: fused LM head forward -> loss assembly -> fused LM head backward.

This is the B200 replacement for the body of ``PolicyWorkerBase._forward_backward_micro``
(skyrl-train/skyrl_train/workers/worker.py:787-900) *after* the transformer trunk has produced the last hidden
states: it consumes ``hidden`` instead of logits, never builds ``[tokens x vocab]``, and returns the same status
dict the reference reports (``final_loss, policy_loss, policy_entropy, policy_kl, loss_metrics/clip_ratio``) plus
the gradients that continue into the trunk (``d_hidden``) and the lm_head (``d_weight``, fp32).

By default one C-ABI call (``skyb200_policy_step``) runs the whole step chunk by chunk (forward GEMM -> loss ->
backward GEMMs, 6*H*V executed flop per token, small kernels on an auxiliary stream); ``fused=False`` issues the
three calls of the autograd path instead.  There is no host synchronisation inside; the only sync is the single
D2H copy of the packed stats vector (skipped with ``sync=False``).  Use one caller thread per device: the step
keeps per-device scratch (auxiliary stream, lockstep counters).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from . import _lib
from .losses import _cfg, _make_params, _off_policy_enabled, compute_off_policy_correction, status_from_packed


@dataclass
class PolicyUpdateOutput:
    status: Dict[str, float]           # python floats (or {"packed": tensor} when sync=False)
    stats: torch.Tensor                # packed fp32 stats vector on the device (SKYB200_STAT_* layout)
    log_probs: torch.Tensor            # [B, A] fp32
    entropy: torch.Tensor              # [B, A] fp32 (masked by entropy_mask if one was given)
    d_hidden: Optional[torch.Tensor]   # [B*A, H] bf16
    d_weight: Optional[torch.Tensor]   # [V, H] fp32 (the buffer passed in, or a new one)
    hidden: Optional[torch.Tensor] = None  # device copy of the hidden states when they were streamed in from the host


class _Workspace:
    """Grow-only per-device scratch so steady-state steps do not touch the allocator."""

    def __init__(self):
        self.bufs: Dict[tuple, torch.Tensor] = {}

    def get(self, key: str, nbytes: int, dev: torch.device) -> torch.Tensor:
        # one buffer per (device, stream): two caller streams on one device must not share scratch
        k = (key, dev.index, torch.cuda.current_stream(dev).cuda_stream)
        b = self.bufs.get(k)
        if b is None or b.numel() < nbytes:
            b = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            self.bufs[k] = b
        return b


_WS = _Workspace()


def policy_update_step(
    hidden: torch.Tensor,
    weight: torch.Tensor,
    labels: torch.Tensor,
    old_log_probs: torch.Tensor,
    advantages: torch.Tensor,
    config,
    loss_mask: Optional[torch.Tensor] = None,
    base_log_probs: Optional[torch.Tensor] = None,
    temperature: float = 1.0,
    entropy_mask: Optional[torch.Tensor] = None,
    d_weight: Optional[torch.Tensor] = None,
    accumulate_d_weight: bool = False,
    need_d_hidden: bool = True,
    need_d_weight: bool = True,
    chunk: int = 0,
    sync: bool = True,
    fused: bool = True,
    cu_seqlens: Optional[torch.Tensor] = None,
    rollout_logprobs: Optional[torch.Tensor] = None,
    timing: bool = False,
    dw_scatter=None,
) -> PolicyUpdateOutput:
    """``hidden`` [B, A, H] (or [B*A, H]) bf16, ``weight`` [V, H] bf16, ``labels`` [B, A] int64,
    ``old_log_probs`` / ``advantages`` / ``loss_mask`` / ``base_log_probs`` [B, A] fp32.

    ``config`` carries the reference's ``AlgorithmConfig`` fields (policy_loss_type, loss_reduction, eps_clip_*,
    clip_ratio_c, use_kl_loss, kl_estimator_type, kl_loss_coef, use_entropy_loss, entropy_loss_coef,
    max_seq_len).  Gradients are those of ``final_loss`` (upstream gradient 1), i.e. what ``loss.backward()``
    delivers in the reference (worker.py:876-877).

    ``hidden`` may also live in (pinned) HOST memory — offloaded activations, or the end-to-end benchmark contract:
    ``skyb200_policy_step_from_host`` then streams it to the GPU that holds ``weight`` one token chunk at a time on a
    copy stream, ahead of the GEMMs, so the transfer hides behind the tensor cores (keep the host tensor alive until
    the step's stream work is done; ``sync=True`` guarantees that).  ``out.hidden`` is the device copy.

    Packed (ragged) layout: pass ``cu_seqlens`` (int32 ``[B + 1]``, ``cu[0] = 0``) and flat per-token tensors —
    ``hidden`` [N, H], ``labels`` / ``old_log_probs`` / ``advantages`` / ``loss_mask`` / ``base_log_probs`` [N] —
    with sequence ``b`` owning tokens ``[cu[b], cu[b+1])``.  Nothing is padded; the result equals the padded layout
    with ``loss_mask = 0`` on the padding (see :func:`pack_by_lengths`).  Outputs are flat [N] then.

    ``fused=True`` (default) runs ``skyb200_policy_step``: per token chunk forward GEMM -> loss -> backward GEMMs,
    the forward keeping a bf16 copy of its logits tiles in a bounded [chunk, V] store so the backward does not
    recompute the projection (6*H*V executed flop per token).  The store is double-buffered: 2 x 18944 tokens x V bf16 =
    11.5 GB of grow-only scratch for a 152k vocabulary (``SKYB200_CHUNK_TILES=37`` halves it at the cost of whole waves
    for the wide backward tiles, ``skyb200_policy_step_workspace`` reports the exact size).  ``fused=False`` runs the three C-ABI calls
    (``head_forward`` / ``policy_loss`` / ``head_backward``) the autograd path uses (8*H*V, fp32 logits in backward).

    ``rollout_logprobs`` [B, A]: with ``config.off_policy_correction`` set (``tis_ratio_type`` /
    ``sequence_mask_metric``) the reference's off-policy correction is applied exactly as the registry losses apply
    it (ppo_utils.py:577-583): TIS weights, outlier-token and sequence masks are computed from ``old_log_probs`` and
    ``rollout_logprobs`` before the step, the loss uses the corrected mask, ``clip_ratio`` keeps the original one,
    and the ``is_ratio_*`` / ``tis_*`` / ``*_sequence_mask_*`` metrics are added to the status dict (padded layout).

    ``dw_scatter`` (a :class:`skyrl_b200.dist.DWeightScatter`, fused only): the data-parallel reduce-scatter of the
    lm_head weight gradient is fused into the ``d_weight`` GEMM - every rank's contributions are averaged straight into
    the owners' ``[V / world, H]`` shards over NVLink while the tensor cores run, no ``[V, H]`` buffer and no
    collective after the step (``out.d_weight`` is None; read ``dw_scatter.shard`` after ``dw_scatter.finish()``).

    ``timing=True`` (fused only) brackets every GEMM launch of the step with CUDA timing events;
    :func:`step_gemm_times` harvests them (live per-kernel durations, what ``bench.py`` reports as ``roofline``).
    """
    lib = _lib.load()
    from_host = not hidden.is_cuda
    if from_host:
        # host-resident activations (offloaded / e2e contract): the C call streams them in chunk by chunk on a copy
        # stream ahead of the GEMMs.  Compute still runs on the GPU that holds `weight` -- there is no CPU fallback.
        if not weight.is_cuda:
            raise RuntimeError("policy_update_step needs the lm_head weight on a CUDA device (no CPU fallback)")
        if not fused:
            raise ValueError("host-resident hidden states are only supported with fused=True")
        if hidden.dtype != torch.bfloat16 or not hidden.is_contiguous():
            raise ValueError("host-resident hidden must be a contiguous bf16 tensor (pin it for an async copy)")
    packed = cu_seqlens is not None
    if packed:
        if not fused:
            raise ValueError("the packed layout is only available with fused=True")
        B = int(cu_seqlens.numel()) - 1
        N = int(old_log_probs.numel())
        A = 1
    else:
        B, A = old_log_probs.shape
        N = B * A
    H = hidden.shape[-1]
    V = weight.shape[0]
    dev = weight.device if from_host else hidden.device
    h = hidden.reshape(N, H)
    if not from_host and (h.dtype != torch.bfloat16 or not h.is_contiguous()):
        h = h.to(torch.bfloat16).contiguous()
    w = weight if (weight.dtype == torch.bfloat16 and weight.is_contiguous()) else weight.to(torch.bfloat16).contiguous()
    lab = labels.reshape(N).to(device=dev, dtype=torch.int64).contiguous()

    def f32(t):
        if t is None:
            return None
        t = t.detach()
        if t.dtype != torch.float32 or t.device != dev:
            t = t.to(device=dev, dtype=torch.float32)
        return t.contiguous()

    olp, adv, mask, base = f32(old_log_probs), f32(advantages), f32(loss_mask), f32(base_log_probs)
    emask = None if entropy_mask is None else f32(entropy_mask).reshape(N)
    loss_type = _cfg(config, "policy_loss_type", "regular")
    if loss_type not in _lib.LOSS_TYPES:
        raise ValueError(f"policy loss '{loss_type}' has no fused B200 kernel; available: {list(_lib.LOSS_TYPES)}")
    use_kl = bool(_cfg(config, "use_kl_loss", False)) and base is not None
    use_ent = bool(_cfg(config, "use_entropy_loss", False))
    reduction = _cfg(config, "loss_reduction", "token_mean")
    if loss_type in ("cross_entropy", "importance_sampling"):
        reduction = "sum"  # Tinker semantics: masked SUM whatever loss_reduction says (ppo_utils.py:892-981)
    if loss_type == "cross_entropy":
        use_kl = use_ent = False  # the SFT path of the worker adds no KL / entropy term (worker.py:810-813)
    # off-policy correction (TIS weights, outlier / sequence masks): depends only on old / rollout log-probs, so it is
    # computed before the step; the same rule as the registry losses (losses._registry_loss)
    tis = clip_mask = None
    opc_metrics: Dict[str, float] = {}
    if loss_type not in ("cross_entropy", "importance_sampling") and _off_policy_enabled(config, rollout_logprobs):
        if packed:
            raise ValueError("the off-policy correction needs the padded [B, A] layout (it reduces per sequence)")
        clip_mask = mask if mask is not None else torch.ones_like(olp)
        tis, opc_metrics, mask = compute_off_policy_correction(olp, f32(rollout_logprobs), mask,
                                                               _cfg(config, "off_policy_correction"))
    params = _make_params(
        loss_type, reduction, config,
        kl_type=_cfg(config, "kl_estimator_type", "k3"),
        kl_coef=float(_cfg(config, "kl_loss_coef", 0.0)) if use_kl else 0.0,
        use_entropy_loss=use_ent, entropy_coef=float(_cfg(config, "entropy_loss_coef", 0.0)),
    )
    T = float(temperature)
    opts = None
    if clip_mask is not None or timing or dw_scatter is not None:
        opts = _lib.StepOpts(tis=_lib.ptr(tis), clip_mask=_lib.ptr(clip_mask), timing=int(bool(timing)))
        if dw_scatter is not None:
            if not fused:
                raise ValueError("dw_scatter (fused reduce-scatter of d_weight) needs fused=True")
            if tuple(dw_scatter.shape) != (V, H):
                raise ValueError(f"dw_scatter was built for a {tuple(dw_scatter.shape)} weight, this head is {(V, H)}")
            opts.dw_scatter = C.pointer(dw_scatter.c_struct)
            need_d_weight = False  # no local [V, H] gradient: the shards are the output
    if fused:
        with torch.cuda.device(dev):
            stream = _lib.stream_ptr(dev)
            logprobs = torch.empty(N, dtype=torch.float32, device=dev)
            ent = torch.empty(N, dtype=torch.float32, device=dev)
            stats = torch.empty(_lib.LOSS_NSTATS, dtype=torch.float32, device=dev)
            need_grad = need_d_hidden or need_d_weight or dw_scatter is not None
            d_hidden = torch.empty((N, H), dtype=torch.bfloat16, device=dev) if need_d_hidden else None
            acc = 0
            if need_d_weight:
                if d_weight is None:
                    d_weight = torch.empty((V, H), dtype=torch.float32, device=dev)
                else:
                    if d_weight.dtype != torch.float32 or tuple(d_weight.shape) != (V, H) or not d_weight.is_contiguous():
                        raise ValueError("d_weight must be a contiguous fp32 [V, H] tensor")
                    acc = 1 if accumulate_d_weight else 0
            cu = None
            if packed:
                cu = cu_seqlens.to(device=dev, dtype=torch.int32).contiguous()
            wsb = lib.skyb200_policy_step_workspace(B, A, N, H, V, params.loss_type, int(need_grad))
            ws = _WS.get("step", wsb, dev)
            tail = (_lib.ptr(w), _lib.ptr(lab), _lib.ptr(olp), _lib.ptr(adv), _lib.ptr(mask),
                    _lib.ptr(base) if use_kl else None, _lib.ptr(emask), _lib.ptr(logprobs), _lib.ptr(ent),
                    _lib.ptr(d_hidden), _lib.ptr(d_weight) if need_d_weight else None, acc, _lib.ptr(stats),
                    _lib.ptr(ws), wsb, B, A, H, V, _lib.ptr(cu), N, T, C.byref(params),
                    C.byref(opts) if opts is not None else None, stream)
            with _lib.nvtx_range("skyb200/policy_step"):
                if from_host:
                    h_host, h = h, torch.empty((N, H), dtype=torch.bfloat16, device=dev)
                    _lib.check(lib.skyb200_policy_step_from_host(h_host.data_ptr(), _lib.ptr(h), *tail),
                               "skyb200_policy_step_from_host")
                else:
                    _lib.check(lib.skyb200_policy_step(_lib.ptr(h), *tail), "skyb200_policy_step")
            if sync:
                status = status_from_packed(stats.tolist(), use_kl)  # the single host sync of the step
                status["response_length"] = A
                for k, v in opc_metrics.items():
                    status["loss_metrics/" + k] = v
            else:
                status = {"packed": stats, **{"loss_metrics/" + k: v for k, v in opc_metrics.items()}}
        if not packed:
            logprobs, ent = logprobs.view(B, A), ent.view(B, A)
        out = PolicyUpdateOutput(status=status, stats=stats, log_probs=logprobs, entropy=ent,
                                 d_hidden=d_hidden, d_weight=d_weight if need_d_weight else None)
        if from_host:
            out.hidden = h  # the device copy of the streamed-in hidden states
        return out
    with torch.cuda.device(dev):
        stream = _lib.stream_ptr(dev)
        logprobs = torch.empty(N, dtype=torch.float32, device=dev)
        ent_raw = torch.empty(N, dtype=torch.float32, device=dev)
        lse2 = torch.empty(N, dtype=torch.float32, device=dev)
        fb = lib.skyb200_head_forward_workspace(N, H, V, chunk)
        ws_f = _WS.get("head_fwd", fb, dev)
        _lib.check(
            lib.skyb200_head_forward(_lib.ptr(h), _lib.ptr(w), _lib.ptr(lab), None, _lib.ptr(logprobs),
                                     _lib.ptr(ent_raw), _lib.ptr(lse2), _lib.ptr(ws_f), fb, N, H, V, T, int(chunk),
                                     stream),
            "skyb200_head_forward",
        )
        ent = ent_raw * emask if emask is not None else ent_raw

        stats = torch.empty(_lib.LOSS_NSTATS, dtype=torch.float32, device=dev)
        d_lp = torch.empty(N, dtype=torch.float32, device=dev)
        d_ent = torch.empty(N, dtype=torch.float32, device=dev) if use_ent else None
        lb = lib.skyb200_policy_loss_workspace(B, A)
        ws_l = _WS.get("loss", lb, dev)
        _lib.check(
            lib.skyb200_policy_loss(_lib.ptr(logprobs), _lib.ptr(olp), _lib.ptr(adv), _lib.ptr(mask),
                                    _lib.ptr(base) if use_kl else None, _lib.ptr(ent), _lib.ptr(d_lp),
                                    _lib.ptr(d_ent), _lib.ptr(stats), _lib.ptr(ws_l), lb, B, A, C.byref(params),
                                    _lib.ptr(tis), _lib.ptr(clip_mask), stream),
            "skyb200_policy_loss",
        )
        d_hidden = None
        if need_d_hidden or need_d_weight:
            g_ent = None
            if use_ent:
                g_ent = d_ent * emask if emask is not None else d_ent
            d_hidden = torch.empty((N, H), dtype=torch.bfloat16, device=dev) if need_d_hidden else None
            acc = 0
            if need_d_weight:
                if d_weight is None:
                    d_weight = torch.empty((V, H), dtype=torch.float32, device=dev)
                else:
                    if d_weight.dtype != torch.float32 or tuple(d_weight.shape) != (V, H) or not d_weight.is_contiguous():
                        raise ValueError("d_weight must be a contiguous fp32 [V, H] tensor")
                    acc = 1 if accumulate_d_weight else 0
            bb = lib.skyb200_head_backward_workspace(N, H, V, chunk)
            ws_b = _WS.get("head_bwd", bb, dev)
            _lib.check(
                lib.skyb200_head_backward(_lib.ptr(h), _lib.ptr(w), _lib.ptr(lab), _lib.ptr(lse2),
                                          _lib.ptr(ent_raw) if g_ent is not None else None, _lib.ptr(d_lp),
                                          _lib.ptr(g_ent), _lib.ptr(d_hidden),
                                          _lib.ptr(d_weight) if need_d_weight else None, acc, _lib.ptr(ws_b), bb, N,
                                          H, V, T, int(chunk), stream),
                "skyb200_head_backward",
            )
        if sync:
            status = status_from_packed(stats.tolist(), use_kl)  # the single host sync of the step
            status["response_length"] = A
            for k, v in opc_metrics.items():
                status["loss_metrics/" + k] = v
        else:
            status = {"packed": stats, **{"loss_metrics/" + k: v for k, v in opc_metrics.items()}}
    return PolicyUpdateOutput(status=status, stats=stats, log_probs=logprobs.view(B, A), entropy=ent.view(B, A),
                              d_hidden=d_hidden, d_weight=d_weight if need_d_weight else None)


def step_gemm_times(device=None) -> Dict[str, Dict[str, float]]:
    """Live per-kernel durations of the fused steps run with ``timing=True`` on ``device`` since the last call:
    ``{"forward" | "d_hidden" | "d_weight": {"ms": summed launch time, "launches": count}}``.  Synchronises."""
    lib = _lib.load()
    out = (C.c_double * 6)()
    with torch.cuda.device(device if device is not None else torch.cuda.current_device()):
        _lib.check(lib.skyb200_policy_step_timing(out), "skyb200_policy_step_timing")
    return {name: {"ms": out[i], "launches": int(out[3 + i])} for i, name in enumerate(("forward", "d_hidden", "d_weight"))}


def pack_by_lengths(lengths: torch.Tensor, *tensors: torch.Tensor):
    """Drop the right padding of ``[B, A, ...]`` tensors: returns ``(cu_seqlens int32 [B + 1], packed tensors)``
    where packed tensor k is ``cat_b tensors[k][b, :lengths[b]]``.  ``None`` entries are passed through."""
    lengths = lengths.to(torch.int64)
    B = lengths.numel()
    cu = torch.zeros(B + 1, dtype=torch.int32, device=lengths.device)
    cu[1:] = torch.cumsum(lengths, 0).to(torch.int32)
    out = []
    for t in tensors:
        if t is None:
            out.append(None)
            continue
        A = t.shape[1]
        keep = (torch.arange(A, device=t.device)[None, :] < lengths.to(t.device)[:, None])
        out.append(t[keep])
    return cu, out


class _FusedPolicyLossFn(torch.autograd.Function):
    """final_loss = f(hidden, weight): the whole fused step as ONE autograd node.

    Forward runs ``policy_update_step`` (which already produces d_hidden / d_weight of the final loss); backward
    hands them to autograd scaled by the upstream gradient, so the trunk's backward continues from ``d_hidden``
    and the optimizer sees ``weight.grad`` — the 6*H*V path behind a ``loss.backward()`` interface."""

    @staticmethod
    def forward(ctx, hidden, weight, labels, old_log_probs, advantages, config, loss_mask, base_log_probs,
                temperature, entropy_mask, cu_seqlens, rollout_logprobs=None):
        need_dh, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        out = policy_update_step(hidden.detach(), weight.detach(), labels, old_log_probs, advantages, config,
                                 loss_mask=loss_mask, base_log_probs=base_log_probs, temperature=temperature,
                                 entropy_mask=entropy_mask, need_d_hidden=need_dh, need_d_weight=need_dw,
                                 sync=False, cu_seqlens=cu_seqlens, rollout_logprobs=rollout_logprobs)
        ctx.save_for_backward(out.d_hidden if out.d_hidden is not None else torch.empty(0, device=hidden.device),
                              out.d_weight if out.d_weight is not None else torch.empty(0, device=hidden.device))
        ctx.shapes = (hidden.shape, hidden.dtype, weight.dtype, need_dh, need_dw)
        stats = out.stats
        ctx.mark_non_differentiable(stats, out.log_probs, out.entropy)
        return stats[_lib.STAT_FINAL_LOSS].clone(), stats, out.log_probs, out.entropy

    @staticmethod
    def backward(ctx, g_loss, _gs, _glp, _gent):
        d_hidden, d_weight = ctx.saved_tensors
        h_shape, h_dtype, w_dtype, need_dh, need_dw = ctx.shapes
        g = g_loss.detach().to(torch.float32)
        gh = gw = None
        if need_dh:
            gh = (d_hidden.float() * g).to(h_dtype).view(h_shape)
        if need_dw:
            gw = (d_weight * g).to(w_dtype)
        return gh, gw, None, None, None, None, None, None, None, None, None, None


def fused_policy_loss(hidden, weight, labels, old_log_probs, advantages, config, loss_mask=None,
                      base_log_probs=None, temperature: float = 1.0, entropy_mask=None, cu_seqlens=None,
                      rollout_logprobs=None):
    """Autograd-facing form of the fused step: returns ``(final_loss, stats, log_probs, entropy)`` where
    ``final_loss`` is a 0-d tensor whose ``backward()`` delivers ``d_hidden`` to the trunk and ``weight.grad``.

    Equivalent to ``fused_lm_head`` + ``policy_update_loss`` + ``loss.backward()`` (and to the reference's
    worker.py:787-877) but with the 6*H*V executed-flop schedule of :func:`policy_update_step`."""
    return _FusedPolicyLossFn.apply(hidden, weight, labels, old_log_probs, advantages, config, loss_mask,
                                    base_log_probs, temperature, entropy_mask, cu_seqlens, rollout_logprobs)
