"""ctypes binding of libskyb200.so — the C ABI declared in include/skyb200.h.

There is deliberately no fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_LIB_PATH = Path(os.environ.get("SKYB200_LIB") or (Path(__file__).resolve().parent / "libskyb200.so"))
_lib = None

# enums (mirror include/skyb200.h)
MASK_F32, MASK_I64, MASK_NONE = 0, 1, 2
LOSS_TYPES = {"regular": 0, "dual_clip": 1, "gspo": 2, "sapo": 3, "cispo": 4, "cross_entropy": 5,
              "importance_sampling": 6}
REDUCTIONS = {"token_mean": 0, "sequence_mean": 1, "seq_mean_token_sum_norm": 2, "sum": 3}
KL_TYPES = {"k1": 0, "abs": 1, "k2": 2, "k3": 3}
LOSS_NSTATS = 16
STAT_POLICY_LOSS, STAT_CLIP_RATIO, STAT_KL_LOSS, STAT_ENTROPY, STAT_FINAL_LOSS, STAT_MASK_SUM = 0, 1, 2, 3, 4, 5
STAT_RATIO_MEAN, STAT_RATIO_MAX, STAT_RATIO_MIN = 6, 7, 8


class LossParams(C.Structure):
    _fields_ = [
        ("loss_type", C.c_int32),
        ("reduction", C.c_int32),
        ("kl_type", C.c_int32),
        ("kl_has_grad", C.c_int32),
        ("use_entropy_loss", C.c_int32),
        ("reserved0", C.c_int32),
        ("eps_clip_low", C.c_float),
        ("eps_clip_high", C.c_float),
        ("clip_ratio_c", C.c_float),
        ("max_seq_len", C.c_float),
        ("kl_loss_coef", C.c_float),
        ("entropy_loss_coef", C.c_float),
        ("sapo_tau_pos", C.c_float),
        ("sapo_tau_neg", C.c_float),
    ]


class OpcParams(C.Structure):
    _fields_ = [
        ("tis_type", C.c_int32),
        ("seq_mask_type", C.c_int32),
        ("has_outlier_low", C.c_int32),
        ("has_outlier_high", C.c_int32),
        ("token_tis_cap", C.c_float),
        ("seq_tis_cap", C.c_float),
        ("outlier_low", C.c_float),
        ("outlier_high", C.c_float),
        ("geo_low", C.c_float),
        ("geo_high", C.c_float),
        ("prod_low", C.c_float),
        ("prod_high", C.c_float),
    ]


class DwScatter(C.Structure):
    """skyb200_dw_scatter: peer-mapped gradient shards of the fused d_weight GEMM + reduce-scatter."""
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("shard", C.c_void_p * 8), ("scale", C.c_float),
                ("remote_red", C.c_int32)]


class StepOpts(C.Structure):
    """skyb200_step_opts: optional inputs of the fused step (off-policy correction, timing, fused reduce-scatter)."""
    _fields_ = [("tis", C.c_void_p), ("clip_mask", C.c_void_p), ("timing", C.c_int64),
                ("dw_scatter", C.POINTER(DwScatter)), ("reserved", C.c_int64 * 4)]


OPC_NMETRICS = 16
LOSS_KL_COV = 7  # skyb200_policy_loss only (global token selection): not in LOSS_TYPES, the fused step rejects it
COV_KL, COV_CLIP = 0, 1


class CovParams(C.Structure):
    _fields_ = [
        ("mode", C.c_int32),
        ("reserved0", C.c_int32),
        ("eps_clip_low", C.c_float),
        ("eps_clip_high", C.c_float),
        ("cov_lb", C.c_float),
        ("cov_ub", C.c_float),
        ("frac", C.c_double),
        ("value_selected", C.c_float),
        ("value_other", C.c_float),
        ("seed", C.c_uint64),
    ]


_P = C.c_void_p
_I64 = C.c_int64
_I32 = C.c_int32
_F = C.c_float
_SZ = C.c_size_t

# name -> (restype, argtypes); every symbol include/skyb200.h declares
SIGNATURES = {
    "skyb200_abi_version": (C.c_int, []),
    "skyb200_last_error": (C.c_char_p, []),
    "skyb200_check_device": (C.c_int, []),
    "skyb200_policy_loss_workspace": (_SZ, [_I64, _I64]),
    "skyb200_policy_loss": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _I64, _I64,
                                      C.POINTER(LossParams), _P, _P, _P]),
    "skyb200_off_policy_workspace": (_SZ, [_I64]),
    "skyb200_off_policy_correction": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _SZ, _I64, _I64, C.POINTER(OpcParams),
                                                _P]),
    "skyb200_value_head_forward": (C.c_int, [_P, _P, _F, _P, _I64, _I64, _P]),
    "skyb200_value_head_backward_workspace": (_SZ, [_I64, _I64]),
    "skyb200_value_head_backward": (C.c_int, [_P, _P, _P, _P, _P, _I32, _P, _SZ, _I64, _I64, _P]),
    "skyb200_scale_inplace": (C.c_int, [_P, _P, _I64, _P]),
    "skyb200_masked_mean": (C.c_int, [_P, _P, _P, _P, _SZ, _I64, _P]),
    "skyb200_approx_kl": (C.c_int, [_P, _P, _P, _P, _I64, _I32, _P]),
    "skyb200_group_advantage": (C.c_int, [_P, _P, _I32, _P, _P, _P, _P, _P, _I64, _I64, _I32, _I32, _F, _P]),
    "skyb200_group_advantage_uniform_supported": (C.c_int, [_I64, _I64]),
    "skyb200_group_advantage_uniform": (C.c_int, [_P, _P, _I32, _P, _P, _P, _I64, _I64, _I32, _I32, _F, _P]),
    "skyb200_gae_workspace": (_SZ, [_I64]),
    "skyb200_gae_advantage": (C.c_int, [_P, _P, _P, _I32, _P, _P, _P, _P, _SZ, _I64, _I64, _F, _F, _P]),
    "skyb200_gae_scan": (C.c_int, [_P, _P, _P, _I32, _P, _P, _P, _P, _SZ, _I64, _I64, _F, _F, _P]),
    "skyb200_whiten_from_moments": (C.c_int, [_P, _P, _P, _I64, _P]),
    "skyb200_reinforce_pp_advantage": (C.c_int, [_P, _P, _I32, _P, _P, _P, _P, _SZ, _I64, _I64, _F, _P]),
    "skyb200_normalize_advantages_workspace": (_SZ, []),
    "skyb200_normalize_advantages": (C.c_int, [_P, _P, _I32, _P, _P, _P, _SZ, _I64, _P]),
    "skyb200_critic_loss": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _SZ, _I64, _I64, _F, _I32, _P]),
    "skyb200_gather_rows": (C.c_int, [_P, _P, _P, _I64, _I64, _I64, _P]),
    "skyb200_reward_kl_penalty": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I32, _F, _P]),
    "skyb200_head_forward_workspace": (_SZ, [_I64, _I64, _I64, _I64]),
    "skyb200_head_backward_workspace": (_SZ, [_I64, _I64, _I64, _I64]),
    "skyb200_head_forward": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _SZ, _I64, _I64, _I64, _F, _I64, _P]),
    "skyb200_cov_select_workspace": (_SZ, [_I64, _I64]),
    "skyb200_cov_select": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _SZ, _I64, _I64, C.POINTER(CovParams), _P]),
    "skyb200_head_forward_stats": (C.c_int, [_P, _P, _P, _P, _P, _SZ, _I64, _I64, _I64, _F, _I64, _P]),
    "skyb200_head_combine_stats": (C.c_int, [_P, _I32, _I64, _F, _P, _P, _P, _P]),
    "skyb200_head_backward": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I32, _P, _SZ, _I64, _I64, _I64, _F,
                                        _I64, _P]),
    "skyb200_policy_step_workspace": (_SZ, [_I64, _I64, _I64, _I64, _I64, _I32, _I32]),
    "skyb200_policy_step_plan": (C.c_int, [_I64, _I64, _I64, _I64, _I64, _I32, _I32, C.POINTER(C.c_int64)]),
    "skyb200_policy_step": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I32, _P, _P, _SZ, _I64, _I64,
                                      _I64, _I64, _P, _I64, _F, C.POINTER(LossParams), C.POINTER(StepOpts), _P]),
    "skyb200_policy_step_from_host": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I32, _P, _P, _SZ,
                                                _I64, _I64, _I64, _I64, _P, _I64, _F, C.POINTER(LossParams),
                                                C.POINTER(StepOpts), _P]),
    "skyb200_policy_step_timing": (C.c_int, [C.POINTER(C.c_double)]),
    "skyb200_dw_group_plan": (C.c_int, [_I64, _I64, _I64]),
    "skyb200_head_tile_plan": (C.c_int, [_I64, _I64, _I64, C.POINTER(C.c_int32)]),
    "skyb200_ipc_export": (C.c_int, [_P, _P, C.POINTER(C.c_int64)]),
    "skyb200_ipc_open": (C.c_int, [_P, _I64, C.POINTER(C.c_void_p)]),
    "skyb200_ipc_close": (C.c_int, [_P]),
    "skyb200_logits_logprob_entropy": (C.c_int, [_P, _I32, _P, _P, _P, _P, _P, _I64, _I64, _P]),
    "skyb200_logits_backward": (C.c_int, [_P, _I32, _P, _P, _P, _P, _P, _P, _I64, _I64, _P]),
    "skyb200_gemm_selftest": (C.c_int, [_P, _P, _P, _I64, _I64, _I64, _I32, _I32, _I32, _I32, _I32, _P]),
    "skyb200_gemm_bench": (C.c_int, [_P, _P, _P, _P, _I64, _I64, _I64, _I32, _I32, _I32, _P]),
    "skyb200_head_gemm_bench": (C.c_int, [_I32, _P, _P, _P, _P, _P, _I32, _I64, _I64, _I64, _P]),
    "skyb200_gemm_debug_record": (C.POINTER(C.c_uint32), []),
}


def lib_path() -> Path:
    return _LIB_PATH


def load():
    """Load libskyb200.so (once).  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        if os.environ.get("SKYB200_AUTOBUILD", "0") == "1":
            from . import build as _build

            _build.build()
        else:
            raise RuntimeError(
                f"{_LIB_PATH} is missing: build it with `python -m skyrl_b200.build` "
                "(there is no CPU or PyTorch fallback for the policy-update kernels)"
            )
    lib = C.CDLL(str(_LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and the header ever diverge
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class Skyb200Error(RuntimeError):
    pass


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().skyb200_last_error()
        raise Skyb200Error(f"{what or 'skyb200 call'} failed (code {rc}): {msg.decode() if msg else ''}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device=None):
    """The current torch stream of ``device`` (default: the current device).  Call it INSIDE the
    ``torch.cuda.device`` context the kernels are launched in, or pass the device: a stream taken under another
    current device belongs to that device."""
    import torch

    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def gemm_debug_record():
    rec = load().skyb200_gemm_debug_record()
    if not rec:
        return None
    names = ["flag", "role", "which", "stage", "parity", "tile", "block", "pad"]
    return {n: int(rec[i]) for i, n in enumerate(names)}


class _NullRange:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def nvtx_range(name: str):
    """``with nvtx_range("skyb200/policy_step"):`` - an NVTX range when ``SKYB200_NVTX=1`` (for
    ``ncu --nvtx --nvtx-include "skyb200/policy_step/"``), otherwise free."""
    if os.environ.get("SKYB200_NVTX", "0") != "1":
        return _NullRange()
    import torch

    return torch.cuda.nvtx.range(name)

