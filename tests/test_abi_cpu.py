"""The C-ABI shared library builds, loads on a CPU-only box and exports every symbol include/skyb200.h declares.
No compute call is made here (no GPU)."""
import ctypes
import os
import re

import pytest

from skyrl_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "skyb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(skyb200_\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    if not _lib.lib_path().exists():
        from skyrl_b200 import build

        build.build()
    return _lib.load()


def test_header_and_binding_agree():
    declared = _declared_symbols()
    assert len(declared) >= 20
    assert sorted(_lib.SIGNATURES) == declared


def test_every_declared_symbol_is_exported(lib):
    raw = ctypes.CDLL(str(_lib.lib_path()))
    for name in _declared_symbols():
        assert hasattr(raw, name), f"{name} declared in skyb200.h but not exported by libskyb200.so"


def test_host_only_entry_points(lib):
    assert lib.skyb200_abi_version() == 1
    assert lib.skyb200_policy_loss_workspace(256, 4096) >= 256 * 8
    assert lib.skyb200_gae_workspace(512) >= 512 * 24
    fwd = lib.skyb200_head_forward_workspace(1 << 20, 3584, 152064, 0)
    bwd = lib.skyb200_head_backward_workspace(1 << 20, 3584, 152064, 0)
    # bounded scratch: independent of N once N exceeds one chunk, and far below N*V*2 bytes (= 319 GB)
    assert fwd == lib.skyb200_head_forward_workspace(1 << 22, 3584, 152064, 0)
    assert bwd == lib.skyb200_head_backward_workspace(1 << 22, 3584, 152064, 0)
    assert bwd < 4 * 2**30 and fwd < 2**30


def test_loss_params_layout_matches_header():
    assert ctypes.sizeof(_lib.LossParams) == 6 * 4 + 8 * 4


def test_bad_arguments_report_errors_without_a_gpu(lib):
    # argument validation happens before any CUDA call
    p = _lib.LossParams()
    rc = lib.skyb200_policy_loss(None, None, None, None, None, None, None, None, None, None, 0, 4, 4,
                                 ctypes.byref(p), None, None, None)
    assert rc == 1
    assert b"null pointer" in lib.skyb200_last_error()
    rc = lib.skyb200_approx_kl(ctypes.c_void_p(16), ctypes.c_void_p(16), None, ctypes.c_void_p(16), 4, 9, None)
    assert rc == 1 and b"Invalid KL estimator type" in lib.skyb200_last_error()
