"""The oracle (oracle/policy_oracle.py) against (1) the golden fixtures generated from the unmodified reference,
(2) the known-answer vectors of the reference's own unit tests, (3) the live reference when /root/reference exists."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import policy_oracle as O
from oracle import ref_loader
from conftest import assert_close

T = torch.tensor


def _t(x):
    return torch.from_numpy(np.asarray(x))


# ---------------------------------------------------------------- golden fixtures (reference outputs)
def test_golden_advantages(golden_adv):
    g = golden_adv
    r, dense, v = _t(g["rewards"]), _t(g["dense_rewards"]), _t(g["values"])
    im = _t(g["mask_i64"])
    fm = im.float()
    uids = g["uids"]
    assert_close(O.grpo_advantage(r, im, uids)[0], g["grpo_std"], what="grpo_std")
    assert_close(O.grpo_advantage(r, fm, uids, norm_by_std=False)[0], g["grpo_nostd"], what="grpo_nostd")
    assert_close(O.grpo_advantage(dense, fm, uids)[0], g["grpo_dense"], what="grpo_dense")
    assert_close(O.rloo_advantage(r, fm, uids)[0], g["rloo"], what="rloo")
    assert_close(O.rloo_advantage(dense, im, uids)[0], g["rloo_dense"], what="rloo_dense")
    for name, (gm, lm) in {"gae_1_095": (1.0, 0.95), "gae_099_1": (0.99, 1.0), "gae_1_1": (1.0, 1.0)}.items():
        adv, ret = O.gae_advantage_return(dense, v, fm, gm, lm)
        assert_close(adv, g[name + "_adv"], what=name + " adv")
        assert_close(ret, g[name + "_ret"], what=name + " ret")


def test_golden_kl(golden_adv):
    g = golden_adv
    lp, base, fm, r = _t(g["lp"]), _t(g["base"]), _t(g["mask_i64"]).float(), _t(g["rewards"])
    for kind in ("k1", "abs", "k2", "k3"):
        assert_close(O.approx_kl(lp, base, fm, kind), g["kl_" + kind], what="kl " + kind)
        assert_close(O.approx_kl(lp, base, None, kind), g["kl_nomask_" + kind], what="kl nomask " + kind)
        out, met = O.reward_kl_penalty(r, lp, base, fm, kind, 0.05)
        assert_close(out, g["klrew_" + kind], what="klrew " + kind)
        assert_close([met["avg_kl"], met["avg_kl_max"]], g["klrew_metrics_" + kind], what="klrew metrics " + kind)


@pytest.mark.parametrize("kind", ["regular", "dual_clip", "gspo"])
@pytest.mark.parametrize("red", ["token_mean", "sequence_mean", "seq_mean_token_sum_norm"])
@pytest.mark.parametrize("mname", ["mask", "nomask"])
def test_golden_losses(golden_loss, kind, red, mname):
    g = golden_loss
    lp, olp, adv, fm = _t(g["lp"]), _t(g["olp"]), _t(g["adv"]), _t(g["mask"])
    cfg = O.make_cfg(policy_loss_type=kind, loss_reduction=red, max_seq_len=64,
                     eps_clip_high=0.28 if kind == "dual_clip" else 0.2)
    x = lp.clone().requires_grad_(True)
    loss, clip = O.policy_loss(x, olp, adv, cfg, fm if mname == "mask" else None)
    loss.backward()
    key = f"{kind}__{red}__{mname}"
    assert_close(loss, g[key + "__loss"], what=key + " loss")
    assert_close(clip, g[key + "__clip"], what=key + " clip")
    assert_close(x.grad, g[key + "__grad"], what=key + " grad")


@pytest.mark.parametrize("name,kw", [
    ("asm_grpo_kl", dict(policy_loss_type="regular", use_kl_loss=True, use_entropy_loss=False)),
    ("asm_ent", dict(policy_loss_type="dual_clip", use_kl_loss=False, use_entropy_loss=True)),
    ("asm_gspo_all", dict(policy_loss_type="gspo", loss_reduction="sequence_mean", use_kl_loss=True,
                          use_entropy_loss=True, kl_estimator_type="k2")),
])
def test_golden_loss_assembly(golden_loss, name, kw):
    g = golden_loss
    lp, olp, adv, fm, base, ent = (_t(g[k]) for k in ("lp", "olp", "adv", "mask", "base", "ent"))
    cfg = O.make_cfg(max_seq_len=64, **kw)
    x = lp.clone().requires_grad_(True)
    e = ent.clone().requires_grad_(True)
    final, st = O.loss_assembly(x, olp, adv, cfg, fm, base, e)
    final.backward()
    want = g[name + "__stats"]
    assert_close(st["final_loss"], want[0], what="final")
    assert_close(st["policy_loss"], want[1], what="policy")
    assert_close(st["policy_entropy"], want[2], what="entropy")
    assert_close(st.get("policy_kl", 0.0), want[3], what="kl")
    assert_close(st["loss_metrics/clip_ratio"], want[4], what="clip")
    assert_close(x.grad, g[name + "__dlp"], what="d_lp")
    assert_close(e.grad if e.grad is not None else torch.zeros_like(e), g[name + "__dent"], what="d_ent")


def test_golden_head(golden_head):
    g = golden_head
    hidden, weight, labels = _t(g["hidden"]), _t(g["weight"]), _t(g["labels"])
    for Tm in (1.0, 0.6):
        k = f"T{Tm}"
        h = hidden.clone().requires_grad_(True)
        w = weight.clone().requires_grad_(True)
        lp, ent = O.head_logprobs_entropy(h, w, labels, Tm)
        assert_close(lp, g[k + "_lp"], rel=1e-5, what=k + " lp")
        assert_close(ent, g[k + "_ent"], rel=1e-5, what=k + " ent")
        g_lp, g_ent = _t(g[k + "_g_lp"]), _t(g[k + "_g_ent"])
        if Tm == 1.0:
            (lp * g_lp).sum().backward()
            assert_close(h.grad, g[k + "_dh_lp"], rel=1e-4, what="dh")
            assert_close(w.grad, g[k + "_dw_lp"], rel=1e-4, what="dw")
        else:
            ((lp * g_lp).sum() + (ent * g_ent).sum()).backward()
            assert_close(h.grad, g[k + "_dh_both"], rel=1e-4, what="dh")
            assert_close(w.grad, g[k + "_dw_both"], rel=1e-4, what="dw")
    assert_close(O.logprobs_from_logits(_t(g["logits"]), _t(g["labels2"])), g["logits_lp"], what="logits lp")
    assert_close(O.entropy_from_logits(_t(g["logits"]), _t(g["amask"])), g["logits_ent"], what="logits ent")


# ---------------------------------------------------------------- known-answer vectors of the reference's tests
def test_kat_kl_estimators():  # tests/cpu/utils/test_ppo_utils.py:27-62
    lp, base, m = T([[0.2, 0.3, 0.5]]), T([[0.1, 0.2, 0.4]]), T([[1.0, 1.0, 0.0]])
    d = lp - base
    assert torch.allclose(O.approx_kl(lp, base, m, "k1"), d * m, atol=1e-4)
    assert torch.allclose(O.approx_kl(lp, base, m, "abs"), d.abs() * m, atol=1e-4)
    assert torch.allclose(O.approx_kl(lp, base, m, "k2"), 0.5 * d * d * m, atol=1e-4)
    x = -d
    assert torch.allclose(O.approx_kl(lp, base, m, "k3"), (x.exp() - 1 - x) * m, atol=1e-4)


def test_kat_gae():  # tests/cpu/utils/test_ppo_utils.py:175-242
    r, v = T([[1.0, 2.0, 3.0]]), T([[0.5, 1.0, 1.5]])
    ones = torch.ones(1, 3)
    _, ret = O.gae_advantage_return(r, v, ones, 1.0, 1.0)
    assert torch.allclose(ret, T([[6.0, 5.0, 3.0]]), atol=1e-5)
    adv, ret = O.gae_advantage_return(r, v, T([[1.0, 0.0, 1.0]]), 1.0, 1.0)  # mask does NOT reset the scan
    assert torch.allclose(ret, T([[6.0, 5.0, 3.0]]), atol=1e-5)
    assert torch.allclose(adv, T([[0.7071, 0.1768, -0.7071]]), atol=1e-4)
    _, ret = O.gae_advantage_return(r, v, ones, 0.5, 1.0)
    assert torch.allclose(ret, T([[2.75, 3.5, 3.0]]), atol=1e-5)
    _, ret = O.gae_advantage_return(r, v, ones, 1.0, 0.5)
    assert torch.allclose(ret, T([[3.625, 4.25, 3.0]]), atol=1e-5)


def test_kat_grpo_rloo():  # tests/cpu/utils/test_ppo_utils.py:103-172
    r = T([[1.0, 2.0, 3.0], [1.0, 1.0, 1.0], [2.0, 3.0, 4.0], [3.0, 4.0, 5.0]])  # sums 6, 3, 9, 12
    m = torch.ones(4, 3)
    adv, ret = O.grpo_advantage(r, m, np.array([0, 0, 1, 1]), norm_by_std=False)
    assert torch.allclose(adv, T([1.5, -1.5, -1.5, 1.5]).unsqueeze(-1) * m, atol=1e-5)
    assert adv is ret
    r5 = torch.cat([r, T([[1.0, 0.0, 0.0]])])
    adv, _ = O.rloo_advantage(r5, torch.ones(5, 3), np.array([0, 0, 1, 1, 2]))
    assert torch.allclose(adv, T([3.0, -3.0, -3.0, 3.0, 0.0]).unsqueeze(-1) * torch.ones(5, 3), atol=1e-5)


def test_kat_reduce_loss():  # tests/cpu/utils/test_ppo_utils.py:245-268
    loss, mask = T([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]]), T([[1.0, 1.0, 1.0], [1.0, 0.0, 0.0]])
    assert math.isclose(O.reduce_loss(loss, mask, "token_mean").item(), 2.5, rel_tol=1e-6)
    assert math.isclose(O.reduce_loss(loss, mask, "sequence_mean").item(), 3.0, rel_tol=1e-6)
    assert math.isclose(O.reduce_loss(loss, mask, "seq_mean_token_sum_norm", 4).item(), 1.25, rel_tol=1e-6)


def test_kat_dual_clip():  # tests/cpu/algorithms/test_losses.py:32-82 -> 4.1667
    adv, olp, lp = T([[1.0, -1.0, -4.0]]), T([[-1.0, -1.0, -3.0]]), T([[-1.69315, -1.0, -0.69741]])
    cfg = O.make_cfg(policy_loss_type="dual_clip", clip_ratio_c=3.0)
    loss, _ = O.policy_loss(lp, olp, adv, cfg, None)
    assert abs(loss.item() - 4.1667) < 1e-4


def test_kat_entropy_uniform():  # tests/cpu/utils/test_torch_utils.py:33-55
    ent = O.entropy_from_logits(torch.zeros(2, 3, 4))
    assert torch.allclose(ent, torch.full((2, 3), math.log(4.0)), atol=1e-4)
    mask = T([[1.0, 0.0, 1.0], [0.0, 1.0, 1.0]])
    assert torch.equal(O.entropy_from_logits(torch.zeros(2, 3, 4), mask) == 0, mask == 0)


def test_masked_var_errors():  # utils/ppo_utils.py:155-160
    with pytest.raises(ValueError):
        O.masked_var(torch.ones(2, 2), torch.zeros(2, 2))
    with pytest.raises(ValueError):
        O.masked_var(torch.ones(2, 2), T([[1.0, 0.0], [0.0, 0.0]]))


# ---------------------------------------------------------------- live reference (build container only)
@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")
def test_oracle_vs_live_reference_random():
    pu, tu = ref_loader.load()
    g = torch.Generator().manual_seed(7)
    B, A = 10, 33
    lens = torch.randint(2, A + 1, (B,), generator=g)
    mask = (torch.arange(A)[None] < lens[:, None]).float()
    r = torch.randn(B, A, generator=g) * mask
    v = torch.randn(B, A, generator=g)
    idx = np.array([f"u{i // 3}" for i in range(B)])
    assert_close(O.grpo_advantage(r, mask, idx)[0], pu.compute_grpo_outcome_advantage(r.clone(), mask, idx)[0])
    a1, r1 = pu.compute_gae_advantage_return(r, v, mask, gamma=0.97, lambd=0.9)
    a2, r2 = O.gae_advantage_return(r, v, mask, 0.97, 0.9)
    assert_close(a2, a1)
    assert_close(r2, r1)
    z = torch.randn(5, 7, 50, generator=g)
    lab = torch.randint(0, 50, (5, 7), generator=g)
    assert_close(O.logprobs_from_logits(z, lab), tu.logprobs_from_logits_v2(z, lab))
    assert_close(O.entropy_from_logits(z), tu.chunked_entropy_from_logits(z))


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_oracle_losses_vs_live_reference_random(seed):
    """Every loss of the registry, oracle vs the unmodified reference on fresh random inputs (value, metric and
    gradient); clip_cov replays the reference's torch.randperm draw from the same global seed."""
    from types import SimpleNamespace as NS

    pu, _ = ref_loader.load()
    g = torch.Generator().manual_seed(100 + seed)
    B, A = 9, 37
    lens = torch.randint(1, A + 1, (B,), generator=g)
    mask = (torch.arange(A)[None] < lens[:, None]).float()
    mask[seed % B] = 0.0
    lp = -torch.rand(B, A, generator=g) * 3
    olp = lp + 0.2 * torch.randn(B, A, generator=g)
    adv = torch.randn(B, A, generator=g)
    extra = dict(sapo=NS(tau_pos=1.0, tau_neg=1.05), cispo=NS(cispo_eps_clip_low=0.2, cispo_eps_clip_high=0.25),
                 clip_cov=NS(clip_ratio=0.2, clip_cov_lb=0.01, clip_cov_ub=4.0),
                 kl_cov=NS(kl_cov_frac=0.3, ppo_kl_coef=0.7), max_seq_len=A,
                 off_policy_correction=NS(tis_ratio_type=None, sequence_mask_metric=None))
    for kind, red in (("regular", "token_mean"), ("dual_clip", "sequence_mean"), ("gspo", "sequence_mean"),
                      ("sapo", "sequence_mean"), ("cispo", "seq_mean_token_sum_norm"), ("clip_cov", "token_mean"),
                      ("kl_cov", "sequence_mean")):
        cfg = O.make_cfg(policy_loss_type=kind, loss_reduction=red, **extra)
        x_ref = lp.clone().requires_grad_(True)
        torch.manual_seed(5 + seed)
        loss_ref, met_ref = pu.PolicyLossRegistry._functions[kind](x_ref, olp, adv, cfg, loss_mask=mask)
        loss_ref.backward()
        x = lp.clone().requires_grad_(True)
        torch.manual_seed(5 + seed)
        if kind in ("regular", "dual_clip", "gspo"):
            loss, clip = O.policy_loss(x, olp, adv, cfg, mask)
            met = {"clip_ratio": clip}
        elif kind in ("clip_cov", "kl_cov"):
            loss, met, _ = O.cov_policy_loss(kind, x, olp, adv, cfg, mask)
        else:
            loss, met = O.extra_policy_loss(kind, x, olp, adv, cfg, mask)
        loss.backward()
        assert_close(loss, loss_ref, what=f"{kind} loss")
        assert_close(met["clip_ratio"], met_ref["clip_ratio"], what=f"{kind} clip_ratio")
        assert_close(x.grad, x_ref.grad, what=f"{kind} grad")


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")
def test_oracle_advantages_vs_live_reference_property():
    """Property test (hypothesis): all four estimators, KL estimators and KL-in-reward arithmetic of the oracle against
    the unmodified reference over drawn shapes, group structures (singletons, shuffled uids, one big group), interior
    mask holes, fully masked rows and gamma / lambda - the inputs the fixed-seed tests do not vary."""
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st

    pu, _ = ref_loader.load()

    @settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
    @given(B=st.integers(1, 9), A=st.integers(1, 40), G=st.integers(1, 5), seed=st.integers(0, 10_000),
           gamma=st.sampled_from([1.0, 0.99, 0.5]), lambd=st.sampled_from([1.0, 0.95, 0.3]),
           shuffle=st.booleans(), hole=st.booleans(), int_mask=st.booleans())
    def run(B, A, G, seed, gamma, lambd, shuffle, hole, int_mask):
        g = torch.Generator().manual_seed(seed)
        lens = torch.randint(1, A + 1, (B,), generator=g)
        mask = (torch.arange(A)[None] < lens[:, None]).float()
        if hole and A > 3:
            mask[:, 1:3] = 0.0
        if B > 2 and seed % 3 == 0:
            mask[0] = 0.0
        m = mask.long() if int_mask else mask
        idx = np.array([f"uid{i // G}" for i in range(B)])
        if shuffle:
            idx = idx[torch.randperm(B, generator=g).numpy()]
        r = torch.randn(B, A, generator=g) * mask
        v = torch.randn(B, A, generator=g)
        assert_close(O.grpo_advantage(r, m, idx)[0], pu.compute_grpo_outcome_advantage(r.clone(), m, idx)[0])
        assert_close(O.grpo_advantage(r, m, idx, norm_by_std=False)[0],
                     pu.compute_grpo_outcome_advantage(r.clone(), m, idx, grpo_norm_by_std=False)[0])
        assert_close(O.rloo_advantage(r, m, idx)[0], pu.compute_rloo_outcome_advantage(r.clone(), m, idx)[0])
        if mask.sum() >= 2:
            a1, r1 = pu.compute_gae_advantage_return(r, v, mask, gamma=gamma, lambd=lambd)
            a2, r2 = O.gae_advantage_return(r, v, mask, gamma, lambd)
            assert_close(r2, r1, what="gae returns")
            assert_close(a2, a1, rel=1e-4, what="gae advantages")  # whitening amplifies fp32 summation-order noise
            a1, r1 = pu.compute_reinforce_plus_plus_outcome_advantage(r, mask, gamma=gamma)
            a2, r2 = O.reinforce_pp_advantage(r, mask, gamma)
            assert_close(r2, r1, what="reinforce++ returns")
            assert_close(a2, a1, rel=1e-4, what="reinforce++ advantages")
        lp, base = -torch.rand(B, A, generator=g) * 3, -torch.rand(B, A, generator=g) * 3
        for kind in ("k1", "abs", "k2", "k3"):
            assert_close(O.approx_kl(lp, base, mask, kind), pu.compute_approx_kl(lp, base, mask, kind))
            assert_close(O.approx_kl(lp, base, None, kind), pu.compute_approx_kl(lp, base, None, kind))

    run()


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")
def test_oracle_losses_vs_live_reference_property():
    """Property test (hypothesis): regular / dual_clip / gspo x the three reductions x {mask, no mask} against the
    unmodified reference - loss, clip ratio and the autograd gradient - with the inputs that hit the special cases:
    log-ratios beyond the +-20 clamp, exact ties of the two surrogates (lp == old_lp), zero and negative advantages,
    fully masked rows."""
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st

    pu, _ = ref_loader.load()

    @settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
    @given(B=st.integers(1, 7), A=st.integers(1, 24), seed=st.integers(0, 10_000),
           kind=st.sampled_from(["regular", "dual_clip", "gspo"]),
           red=st.sampled_from(["token_mean", "sequence_mean", "seq_mean_token_sum_norm"]),
           use_mask=st.booleans(), eps=st.sampled_from([(0.2, 0.2), (0.2, 0.28), (3e-4, 4e-4)]),
           c=st.sampled_from([3.0, 10.0]))
    def run(B, A, seed, kind, red, use_mask, eps, c):
        g = torch.Generator().manual_seed(seed)
        lens = torch.randint(1, A + 1, (B,), generator=g)
        mask = (torch.arange(A)[None] < lens[:, None]).float()
        if B > 1:
            mask[seed % B] = 0.0
        lp = -torch.rand(B, A, generator=g) * 3
        olp = lp + 0.3 * torch.randn(B, A, generator=g)
        tie = torch.rand(B, A, generator=g) < 0.3
        olp = torch.where(tie, lp, olp)                                   # exact ties of the two surrogates
        if kind != "gspo":
            far = torch.rand(B, A, generator=g) < 0.1
            olp = torch.where(far, lp + 25.0 * torch.sign(torch.randn(B, A, generator=g)), olp)  # beyond the clamp
        adv = torch.randn(B, A, generator=g)
        adv = torch.where(torch.rand(B, A, generator=g) < 0.1, torch.zeros(()), adv)
        cfg = O.make_cfg(policy_loss_type=kind, loss_reduction=red, eps_clip_low=eps[0], eps_clip_high=eps[1],
                         clip_ratio_c=c, max_seq_len=A + 3)
        m = mask if use_mask else None
        x_ref = lp.clone().requires_grad_(True)
        loss_ref, met_ref = pu.PolicyLossRegistry._functions[kind](x_ref, olp, adv, cfg, loss_mask=m)
        loss_ref.backward()
        x = lp.clone().requires_grad_(True)
        loss, clip = O.policy_loss(x, olp, adv, cfg, m)
        loss.backward()
        assert_close(loss, loss_ref, what=f"{kind}/{red} loss")
        assert clip == pytest.approx(met_ref["clip_ratio"], abs=1e-7)
        assert_close(x.grad, x_ref.grad, what=f"{kind}/{red} grad")

    run()


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")
def test_oracle_off_policy_correction_vs_live_reference_property():
    """Property test (hypothesis): TIS (token / sequence), outlier-token thresholds and the geometric / product
    sequence masks of the oracle against the unmodified reference (off_policy_correction_utils.py:186-260) over drawn
    configurations - TIS weights, the updated loss mask and every metric the reference reports."""
    import importlib

    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st
    from types import SimpleNamespace as NS

    ref_loader.load()
    opc_ref = importlib.import_module("skyrl_train.utils.off_policy_correction_utils")

    @settings(max_examples=60, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
    @given(B=st.integers(1, 6), A=st.integers(1, 20), seed=st.integers(0, 10_000),
           tis=st.sampled_from([None, "token", "sequence"]), seqm=st.sampled_from([None, "geometric", "product"]),
           out_lo=st.sampled_from([None, 0.3, 0.8]), out_hi=st.sampled_from([None, 1.5, 4.0]),
           spread=st.sampled_from([0.05, 0.3, 1.0]))
    def run(B, A, seed, tis, seqm, out_lo, out_hi, spread):
        g = torch.Generator().manual_seed(seed)
        lens = torch.randint(1, A + 1, (B,), generator=g)
        mask = (torch.arange(A)[None] < lens[:, None]).float()
        olp = -torch.rand(B, A, generator=g) * 3
        rlp = olp + spread * torch.randn(B, A, generator=g)
        cfg = NS(tis_ratio_type=tis, sequence_mask_metric=seqm, token_tis_ratio_clip_high=2.0,
                 sequence_tis_ratio_clip_high=5.0, outlier_token_is_threshold_low=out_lo,
                 outlier_token_is_threshold_high=out_hi, geo_mask_low=0.9, geo_mask_high=1.1,
                 product_mask_low=0.5, product_mask_high=2.0)
        t_ref, m_ref, k_ref = opc_ref.compute_off_policy_correction(olp, rlp, mask.clone(), cfg)
        t, m, k = O.off_policy_correction(olp, rlp, mask.clone(), cfg)
        assert (t is None) == (t_ref is None)
        if t is not None:
            assert_close(t.expand_as(olp) if t.shape != olp.shape else t,
                         t_ref.expand_as(olp) if t_ref.shape != olp.shape else t_ref, what="tis weights")
        assert torch.equal(k, k_ref), "updated loss mask"
        assert sorted(m) == sorted(m_ref), (sorted(m), sorted(m_ref))
        for key, v in m_ref.items():
            if not (math.isnan(v) and math.isnan(m[key])):     # std of a single element is NaN on both sides
                assert m[key] == pytest.approx(v, rel=1e-5, abs=1e-6), key

    run()


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")
def test_oracle_critic_loss_and_normalisation_vs_live_reference_property():
    """Property test (hypothesis): the clipped value loss (value, clip fraction, gradient; exact ties of the two
    surrogates included) and normalize_advantages_dict of the oracle against the unmodified reference."""
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st
    from types import SimpleNamespace as NS

    pu, _ = ref_loader.load()

    @settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
    @given(B=st.integers(1, 7), A=st.integers(1, 24), seed=st.integers(0, 10_000),
           clip=st.sampled_from([None, 0.2, 0.05]), use_mask=st.booleans())
    def run(B, A, seed, clip, use_mask):
        g = torch.Generator().manual_seed(seed)
        lens = torch.randint(1, A + 1, (B,), generator=g)
        mask = (torch.arange(A)[None] < lens[:, None]).float()
        values = torch.randn(B, A, generator=g)
        old = values + 0.15 * torch.randn(B, A, generator=g)
        old = torch.where(torch.rand(B, A, generator=g) < 0.3, values, old)      # ties: surr1 == surr2
        returns = torch.randn(B, A, generator=g)
        m = mask if use_mask else None
        x_ref = values.clone().requires_grad_(True)
        loss_ref, cf_ref = pu.ppo_critic_loss(x_ref, old, returns, NS(value_clip=clip), loss_mask=m)
        loss_ref.backward()
        x = values.clone().requires_grad_(True)
        loss, cf = O.critic_loss(x, old, returns, clip, m)
        loss.backward()
        assert_close(loss, loss_ref, what="critic loss")
        assert (cf is None) == (cf_ref is None) and (cf is None or cf == pytest.approx(cf_ref, abs=1e-7))
        assert_close(x.grad, x_ref.grad, what="critic gradient")
        adv = torch.randn(B, A, generator=g) * 3 + 1
        want = pu.normalize_advantages_dict({"advantages": adv.clone(), "response_mask": mask})["advantages"]
        assert_close(O.normalize_advantages(adv, mask), want, what="normalize_advantages_dict")

    run()


# ---------------------------------------------------------------- "next" rows
def test_golden_reinforce_pp_and_normalize(golden_adv):
    g = golden_adv
    dense, fm = _t(g["dense_rewards"]), _t(g["mask_i64"]).float()
    for gm in (1.0, 0.9):
        adv, ret = O.reinforce_pp_advantage(dense, fm, gm)
        assert_close(ret, g[f"rpp_{gm}_ret"], what="rpp ret")
        assert_close(adv, g[f"rpp_{gm}_adv"], what="rpp adv")
    adv, ret = O.reinforce_pp_advantage(dense, _t(g["rpp_hole_mask"]), 0.95)
    assert_close(ret, g["rpp_hole_ret"])
    assert_close(adv, g["rpp_hole_adv"])
    assert_close(O.normalize_advantages(_t(g["norm_in"]), fm), g["norm_out"])


def test_kat_reinforce_pp():  # tests/cpu/utils/test_ppo_utils.py:65-100
    r = T([[1.0, 2.0, 3.0]])
    _, ret = O.reinforce_pp_advantage(r, T([[1.0, 1.0, 0.0]]), 1.0)
    assert torch.allclose(ret, T([[3.0, 2.0, 3.0]]), atol=1e-5)  # reset after EOS
    _, ret = O.reinforce_pp_advantage(r, torch.ones(1, 3), 0.5)
    assert torch.allclose(ret, T([[2.75, 3.5, 3.0]]), atol=1e-5)


@pytest.mark.parametrize("cname,clip", [("clip02", 0.2), ("noclip", None)])
@pytest.mark.parametrize("mname", ["mask", "nomask"])
def test_golden_critic_loss(golden_loss, cname, clip, mname):
    g = golden_loss
    v = _t(g["values"]).clone().requires_grad_(True)
    loss, clipfrac = O.critic_loss(v, _t(g["old_values"]), _t(g["returns"]), clip,
                                   _t(g["mask"]) if mname == "mask" else None)
    loss.backward()
    key = f"critic__{cname}__{mname}"
    assert_close(loss, g[key + "__loss"])
    assert_close(-1.0 if clipfrac is None else clipfrac, g[key + "__clipfrac"])
    assert_close(v.grad, g[key + "__grad"])


_XCFG = dict(sapo=__import__("types").SimpleNamespace(tau_pos=1.0, tau_neg=2.0),
             cispo=__import__("types").SimpleNamespace(cispo_eps_clip_low=0.1, cispo_eps_clip_high=0.3))
_XCASES = [("sapo", "sequence_mean"), ("sapo", "token_mean"), ("cispo", "token_mean"),
           ("cispo", "seq_mean_token_sum_norm"), ("cross_entropy", "token_mean"), ("importance_sampling", "token_mean")]


@pytest.mark.parametrize("kind,red", _XCASES)
@pytest.mark.parametrize("mname", ["mask", "nomask"])
def test_golden_extra_losses(golden_loss, kind, red, mname):
    g = golden_loss
    lp, adv, fm = _t(g["lp"]), _t(g["adv"]), _t(g["mask"])
    olp = _t(g["olp_mild"]) if kind == "importance_sampling" else _t(g["olp"])
    cfg = O.make_cfg(policy_loss_type=kind, loss_reduction=red, max_seq_len=64, **_XCFG)
    x = lp.clone().requires_grad_(True)
    loss, met = O.extra_policy_loss(kind, x, olp, adv, cfg, fm if mname == "mask" else None)
    loss.backward()
    key = f"{kind}__{red}__{mname}"
    assert_close(loss, g[key + "__loss"], what="loss")
    assert_close(list(met.values())[0], g[key + "__metric"], what="metric")
    assert_close(x.grad, g[key + "__grad"], what="grad")


_COV_CFG = dict(clip_cov=__import__("types").SimpleNamespace(clip_ratio=0.1, clip_cov_lb=0.05, clip_cov_ub=5.0),
                kl_cov=__import__("types").SimpleNamespace(kl_cov_frac=0.2, ppo_kl_coef=1.0))
_COV_CASES = [("clip_cov", "token_mean"), ("clip_cov", "sequence_mean"), ("kl_cov", "token_mean"),
              ("kl_cov", "seq_mean_token_sum_norm")]


@pytest.mark.parametrize("kind,red", _COV_CASES)
def test_golden_cov_losses(golden_loss, kind, red):
    """clip_cov / kl_cov against the reference; clip_cov's torch.randperm draw is replayed from the same seed."""
    g = golden_loss
    lp, adv, fm, olp = _t(g["lp"]), _t(g["adv"]), _t(g["mask"]), _t(g["olp_mild"])
    cfg = O.make_cfg(policy_loss_type=kind, loss_reduction=red, max_seq_len=64, **_COV_CFG)
    x = lp.clone().requires_grad_(True)
    torch.manual_seed(777)
    loss, met, sel = O.cov_policy_loss(kind, x, olp, adv, cfg, fm)
    loss.backward()
    key = f"{kind}__{red}__mask"
    assert_close(loss, g[key + "__loss"], what="loss")
    assert_close(met["clip_ratio"], g[key + "__metric"], what="metric")
    assert_close(x.grad, g[key + "__grad"], what="grad")
    assert sel.sum().item() >= 1 and torch.all(sel[fm == 0] == 0)


def test_kat_cispo():  # tests/cpu/algorithms/test_losses.py:85-137 -> -0.99768 (cispo eps 0.2 / 0.2 there)
    from types import SimpleNamespace

    adv, olp, lp = T([[1.0, -1.0, -4.0]]), T([[-1.0, -1.0, -3.0]]), T([[-1.69315, -1.0, -0.69741]])
    cfg = O.make_cfg(policy_loss_type="cispo", cispo=SimpleNamespace(cispo_eps_clip_low=0.2, cispo_eps_clip_high=0.2))
    loss, _ = O.extra_policy_loss("cispo", lp, olp, adv, cfg, None)
    assert abs(loss.item() - (-0.99768)) < 2e-3


def _opc_cases(g):
    import json

    for key in sorted(k[:-6] for k in g if k.startswith("opc__") and k.endswith("__loss")):
        yield key, json.loads(str(g[key + "__cfg"])), json.loads(str(g[key + "__metrics"]))


def test_golden_off_policy_correction(golden_loss):
    from types import SimpleNamespace

    g = golden_loss
    lp, olp, adv, fm, rlp = (_t(g[k]) for k in ("lp", "olp", "adv", "mask", "rlp"))
    n = 0
    for key, okw, want_met in _opc_cases(g):
        kind = key.split("__")[-1]
        cfg = O.make_cfg(policy_loss_type=kind, loss_reduction="token_mean" if kind == "regular" else "sequence_mean",
                         eps_clip_high=0.28 if kind == "dual_clip" else 0.2, max_seq_len=64,
                         off_policy_correction=SimpleNamespace(**okw))
        x = lp.clone().requires_grad_(True)
        loss, met = O.policy_loss_with_correction(x, olp, adv, cfg, fm, rlp)
        loss.backward()
        assert_close(loss, g[key + "__loss"], what=key)
        assert_close(x.grad, g[key + "__grad"], what=key + " grad")
        assert sorted(met) == sorted(want_met), key
        for k in want_met:
            assert met[k] == pytest.approx(want_met[k], rel=1e-5, abs=1e-6), (key, k)
        n += 1
    assert n == 8


def test_kat_off_policy_correction():  # tests/cpu/algorithms/test_off_policy_correction.py:20-330
    import opc_kat

    for name, olp, rlp, mask, kw, want_tis, want_keep, want_met in opc_kat.CASES:
        olp_t, rlp_t, m_t = T(olp), T(rlp), T(mask)
        tis, met, new_mask = O.off_policy_correction(olp_t, rlp_t, m_t, opc_kat.cfg(**kw))
        assert torch.allclose(new_mask, m_t * T(want_keep).unsqueeze(-1), atol=1e-6), name
        if want_tis is not None:
            assert torch.allclose(tis.expand_as(olp_t), T(want_tis), rtol=1e-3, atol=1e-4), name
        for k, v in want_met.items():
            assert abs(met[k] - v) < 0.01, (name, k, met[k], v)


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")
def test_committed_goldens_are_what_the_reference_produces_today(tmp_path, monkeypatch):
    """tests/golden/*.npz regenerated with the committed script from the unmodified reference: bit-identical, so the
    fixtures that travel to the GPU box are neither stale nor hand-edited."""
    from oracle import make_golden as mg

    real = mg.OUT
    monkeypatch.setattr(mg, "OUT", str(tmp_path))
    mg.main()
    for name in ("advantages.npz", "losses.npz", "head.npz"):
        old, new = dict(np.load(os.path.join(real, name))), dict(np.load(os.path.join(str(tmp_path), name)))
        assert sorted(old) == sorted(new), name
        for k in old:
            same = np.array_equal(old[k], new[k], equal_nan=old[k].dtype.kind == "f")
            assert old[k].dtype == new[k].dtype and same, (name, k)
