"""The device source of the per-token loss arithmetic (skyrl_b200/csrc/loss_token.cuh - the code loss_main_kernel and
loss_flat_kernel compile) built for the HOST with g++ and checked against the oracle: every loss type in its
compile-time-typed, run-time-typed (-1) and clipped-family (-2) form, the three reductions (+ sum), every KL
estimator with and without its gradient, on inputs with clamp hits, exact surrogate ties, zero advantages and a fully
masked row.  Test infrastructure: nothing in the product loads this build; the GPU tests remain the parity tests of
the compiled kernels."""
import ctypes as C
import os
import shutil
import subprocess
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

from oracle import policy_oracle as O
from skyrl_b200 import _lib
from conftest import assert_close

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    so = tmp_path_factory.mktemp("host") / "loss_token_host.so"
    cmd = [gxx, "-O1", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "host", "loss_token_host.cpp"), "-o", str(so)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr  # -Werror: "used uninitialized" in the device source fails here
    lib = C.CDLL(str(so))
    fp, dp = C.POINTER(C.c_float), C.POINTER(C.c_double)
    lib.skyb200_host_loss_tokens.restype = C.c_int
    lib.skyb200_host_loss_tokens.argtypes = [C.POINTER(_lib.LossParams), C.c_int, C.c_int, C.c_int, fp, fp, fp, fp, fp,
                                             fp, fp, dp]
    return lib


def _inputs(seed, B=6, A=29, gspo=False):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(1, A + 1, (B,), generator=g)
    mask = (torch.arange(A)[None] < lens[:, None]).float()
    mask[seed % B] = 0.0
    lp = -torch.rand(B, A, generator=g) * 3
    olp = lp + 0.3 * torch.randn(B, A, generator=g)
    olp = torch.where(torch.rand(B, A, generator=g) < 0.3, lp, olp)           # exact ties
    if not gspo:
        far = torch.rand(B, A, generator=g) < 0.1
        olp = torch.where(far, lp + 25.0 * torch.sign(torch.randn(B, A, generator=g)), olp)
    base = lp + 0.3 * torch.randn(B, A, generator=g)
    base[0, :3] = lp[0, :3] + torch.tensor([25.0, -25.0, 2.6])
    adv = torch.randn(B, A, generator=g)
    adv = torch.where(torch.rand(B, A, generator=g) < 0.1, torch.zeros(()), adv)
    return lp, olp, adv, mask, base


def _call(lib, params, ktype, lp, olp, adv, mask, base, tis=None):
    B, A = lp.shape
    arr = lambda t: None if t is None else np.ascontiguousarray(t.detach().numpy(), dtype=np.float32)
    ptr = lambda a: None if a is None else a.ctypes.data_as(C.POINTER(C.c_float))
    a = [arr(x) for x in (lp, olp, adv, mask, base, tis)]
    d_lp = np.zeros((B, A), dtype=np.float32)
    out = np.zeros(8, dtype=np.float64)
    rc = lib.skyb200_host_loss_tokens(C.byref(params), ktype, B, A, *[ptr(x) for x in a], ptr(d_lp),
                                      out.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == 0
    return torch.from_numpy(d_lp), out


def _params(kind, red, cfg, **kw):
    from skyrl_b200.losses import _make_params

    return _make_params(kind, red, cfg, **kw)


CLIPPED = ["regular", "dual_clip", "gspo"]
EXTRA = dict(sapo=NS(tau_pos=1.0, tau_neg=1.05), cispo=NS(cispo_eps_clip_low=0.2, cispo_eps_clip_high=0.25))


@pytest.mark.parametrize("red", ["token_mean", "sequence_mean", "seq_mean_token_sum_norm"])
@pytest.mark.parametrize("kind", CLIPPED + ["sapo", "cispo"])
@pytest.mark.parametrize("use_mask", [True, False])
def test_loss_value_clip_ratio_and_gradient(host, kind, red, use_mask):
    for seed in (0, 1):
        lp, olp, adv, mask, _ = _inputs(seed, gspo=(kind == "gspo"))
        m = mask if use_mask else None
        cfg = O.make_cfg(policy_loss_type=kind, loss_reduction=red, eps_clip_low=0.2, eps_clip_high=0.28,
                         clip_ratio_c=3.0, max_seq_len=40, **EXTRA)
        x = lp.clone().requires_grad_(True)
        if kind in CLIPPED:
            ref, clip = O.policy_loss(x, olp, adv, cfg, m)
        else:
            ref, met = O.extra_policy_loss(kind, x, olp, adv, cfg, m)
            clip = met["clip_ratio"]
        ref.backward()
        forms = [_lib.LOSS_TYPES[kind], -1] + ([-2] if kind in CLIPPED else [])
        for ktype in forms:
            d_lp, out = _call(host, _params(kind, red, cfg), ktype, lp, olp, adv, m, None)
            assert_close(out[0], ref.item(), what=f"{kind}/{red} loss (form {ktype})")
            assert out[1] == pytest.approx(clip, abs=1e-6), (kind, ktype)
            assert_close(d_lp, x.grad, floor=1e-3, what=f"{kind}/{red} gradient (form {ktype})")


@pytest.mark.parametrize("kind", ["cross_entropy", "importance_sampling"])
def test_tinker_losses_are_masked_sums(host, kind):
    lp, olp, adv, mask, _ = _inputs(3)
    olp = lp + 0.3 * torch.randn(lp.shape, generator=torch.Generator().manual_seed(5))   # un-clamped ratio: stay finite
    cfg = O.make_cfg(policy_loss_type=kind, loss_reduction="token_mean")
    x = lp.clone().requires_grad_(True)
    ref, met = O.extra_policy_loss(kind, x, olp, adv, cfg, mask)
    ref.backward()
    for ktype in (_lib.LOSS_TYPES[kind], -1):
        d_lp, out = _call(host, _params(kind, "sum", cfg), ktype, lp, olp, adv, mask, None)
        assert_close(out[0], ref.item(), what=kind)
        assert_close(d_lp, x.grad, floor=1e-3, what=kind + " gradient")
        if kind == "importance_sampling":
            assert out[3] == pytest.approx(met["importance_ratio"], rel=1e-5)


@pytest.mark.parametrize("kl", ["k1", "abs", "k2", "k3"])
@pytest.mark.parametrize("has_grad", [False, True])
def test_kl_term_value_and_optional_gradient(host, kl, has_grad):
    lp, olp, adv, mask, base = _inputs(7)
    coef = 0.3
    cfg = O.make_cfg(policy_loss_type="regular", loss_reduction="sequence_mean", use_kl_loss=True,
                     kl_estimator_type=kl, kl_loss_coef=coef)
    x = lp.clone().requires_grad_(True)
    pl, _ = O.policy_loss(x, olp, adv, cfg, mask)
    d = x - base
    if kl == "k3":
        z = (-d).clamp(-20.0, 20.0)
        k = (z.exp() - z - 1.0).clamp(-10.0, 10.0)
    else:
        k = {"k1": d, "abs": d.abs(), "k2": 0.5 * d * d}[kl]
    klm = O.masked_mean(k * mask, mask, dim=-1).mean()
    assert klm.item() == pytest.approx(O.masked_mean(O.approx_kl(lp, base, mask, kl), mask, dim=-1).mean().item(), rel=1e-6)
    (pl + coef * klm if has_grad else pl).backward()
    p = _params("regular", "sequence_mean", cfg, kl_type=kl, kl_coef=coef, kl_has_grad=has_grad)
    for ktype in (0, -1, -2):
        d_lp, out = _call(host, p, ktype, lp, olp, adv, mask, base)
        assert out[2] == pytest.approx(klm.item(), rel=1e-5, abs=1e-6), (kl, ktype)
        assert_close(d_lp, x.grad, floor=1e-3, what=f"gradient with the {kl} KL term (has_grad={has_grad}, form {ktype})")


def test_kl_cov_token_weights(host):
    lp, olp, adv, mask, _ = _inputs(11)
    olp = lp + 0.3 * torch.randn(lp.shape, generator=torch.Generator().manual_seed(6))
    cfg = O.make_cfg(policy_loss_type="kl_cov", loss_reduction="token_mean", kl_cov=NS(kl_cov_frac=0.2, ppo_kl_coef=0.7))
    sel = (torch.rand(lp.shape, generator=torch.Generator().manual_seed(8)) < 0.2).float() * mask
    x = lp.clone().requires_grad_(True)
    ref, _, _ = O.cov_policy_loss("kl_cov", x, olp, adv, cfg, mask, sel=sel)
    ref.backward()
    p = _lib.LossParams()
    p.loss_type, p.reduction = _lib.LOSS_KL_COV, _lib.REDUCTIONS["token_mean"]
    d_lp, out = _call(host, p, _lib.LOSS_KL_COV, lp, olp, adv, mask, None, tis=sel * 0.7)
    assert_close(out[0], ref.item(), what="kl_cov loss")
    assert_close(d_lp, x.grad, floor=1e-3, what="kl_cov gradient")
