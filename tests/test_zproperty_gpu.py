"""Property tests (hypothesis) of the CUDA path against the oracle: the drawn shapes, group structures, masks and
special values of the CPU property tests that pin the oracle to the reference (tests/test_oracle_cpu.py), replayed on
the GPU.  First run on a B200 in round 2 (profiles/r02a_gpu_tests.log) triaged two findings: `is_ratio_std` of a
single element is NaN in the reference (torch.std) and was 0 here - fixed in offpolicy.cu; a suffix scan's fp32 error
scales with its largest running magnitude, not with the cancelled value at a position - the floor below."""
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

import skyrl_b200
from oracle import policy_oracle as O
from conftest import assert_close

pytestmark = [pytest.mark.gpu]
DEV = "cuda:0"


def _settings(n=40):
    from hypothesis import HealthCheck, settings

    return settings(max_examples=n, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)


def test_advantages_property():
    from hypothesis import given
    from hypothesis import strategies as st

    @_settings()
    @given(B=st.integers(1, 12), A=st.sampled_from([1, 3, 17, 64, 257, 1024, 4100]), G=st.integers(1, 5),
           seed=st.integers(0, 10_000), gamma=st.sampled_from([1.0, 0.99, 0.5]), lambd=st.sampled_from([1.0, 0.95, 0.3]),
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
        for name, want in (("grpo", O.grpo_advantage(r, m, idx)[0]), ("rloo", O.rloo_advantage(r, m, idx)[0])):
            got, ret = skyrl_b200.AdvantageEstimatorRegistry.get(name)(token_level_rewards=r.to(DEV),
                                                                       response_mask=m.to(DEV), index=idx)
            assert_close(got, want, what=name)
            assert torch.equal(got.cpu() == 0, want == 0), name + " zero pattern"
        got, _ = skyrl_b200.compute_grpo_outcome_advantage(r.to(DEV), m.to(DEV), idx, grpo_norm_by_std=False)
        assert_close(got, O.grpo_advantage(r, m, idx, norm_by_std=False)[0], what="grpo without std")
        if mask.sum() >= 2:
            a_ref, r_ref = O.gae_advantage_return(r, v, mask, gamma, lambd)
            a, ret = skyrl_b200.compute_gae_advantage_return(r.to(DEV), v.to(DEV), m.to(DEV), gamma, lambd)
            # a suffix scan's fp32 rounding error scales with the largest running magnitude of the scan (the
            # reference's own sequential fp32 sum carries it), not with the possibly cancelled value at a position
            assert_close(ret, r_ref, floor=max(1.0, float(r_ref.abs().max())), what="gae returns")
            assert_close(a, a_ref, rel=1e-4, what="gae advantages")
            a_ref, r_ref = O.reinforce_pp_advantage(r, mask, gamma)
            a, ret = skyrl_b200.compute_reinforce_plus_plus_outcome_advantage(r.to(DEV), m.to(DEV), gamma=gamma)
            assert_close(ret, r_ref, floor=max(1.0, float(r_ref.abs().max())), what="reinforce++ returns")
            assert_close(a, a_ref, rel=1e-4, what="reinforce++ advantages")
        lp, base = -torch.rand(B, A, generator=g) * 3, -torch.rand(B, A, generator=g) * 3
        for kind in ("k1", "abs", "k2", "k3"):
            assert_close(skyrl_b200.compute_approx_kl(lp.to(DEV), base.to(DEV), mask.to(DEV), kind),
                         O.approx_kl(lp, base, mask, kind), what="kl " + kind)

    run()


def test_losses_property():
    from hypothesis import given
    from hypothesis import strategies as st

    @_settings()
    @given(B=st.integers(1, 9), A=st.sampled_from([1, 5, 24, 300, 1024, 4099]), seed=st.integers(0, 10_000),
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
        olp = torch.where(torch.rand(B, A, generator=g) < 0.3, lp, olp)
        if kind != "gspo":
            far = torch.rand(B, A, generator=g) < 0.1
            olp = torch.where(far, lp + 25.0 * torch.sign(torch.randn(B, A, generator=g)), olp)
        adv = torch.randn(B, A, generator=g)
        adv = torch.where(torch.rand(B, A, generator=g) < 0.1, torch.zeros(()), adv)
        cfg = O.make_cfg(policy_loss_type=kind, loss_reduction=red, eps_clip_low=eps[0], eps_clip_high=eps[1],
                         clip_ratio_c=c, max_seq_len=A + 3)
        m = mask if use_mask else None
        x = lp.clone().requires_grad_(True)
        ref, clip = O.policy_loss(x, olp, adv, cfg, m)
        ref.backward()
        y = lp.to(DEV).requires_grad_(True)
        loss, met = skyrl_b200.PolicyLossRegistry.get(kind)(y, olp.to(DEV), adv.to(DEV), cfg,
                                                             loss_mask=None if m is None else m.to(DEV))
        loss.backward()
        assert_close(loss, ref, what=f"{kind}/{red} loss")
        assert met["clip_ratio"] == pytest.approx(clip, abs=1e-6)
        assert_close(y.grad, x.grad, floor=1e-3, what=f"{kind}/{red} gradient")

    run()


def test_off_policy_correction_property():
    from hypothesis import given
    from hypothesis import strategies as st

    @_settings(60)
    @given(B=st.integers(1, 8), A=st.sampled_from([1, 7, 100, 1500]), seed=st.integers(0, 10_000),
           tis=st.sampled_from([None, "token", "sequence"]), seqm=st.sampled_from([None, "geometric", "product"]),
           out_lo=st.sampled_from([None, 0.3, 0.8]), out_hi=st.sampled_from([None, 1.5, 4.0]),
           spread=st.sampled_from([0.05, 0.3]))
    def run(B, A, seed, tis, seqm, out_lo, out_hi, spread):
        g = torch.Generator().manual_seed(seed)
        lens = torch.randint(1, A + 1, (B,), generator=g)
        mask = (torch.arange(A)[None] < lens[:, None]).float()
        lp = -torch.rand(B, A, generator=g) * 3
        olp = lp + 0.1 * torch.randn(B, A, generator=g)
        rlp = olp + spread * torch.randn(B, A, generator=g) / max(1.0, A ** 0.5 / 4)   # keep sequence products finite
        adv = torch.randn(B, A, generator=g)
        opc = NS(tis_ratio_type=tis, sequence_mask_metric=seqm, token_tis_ratio_clip_high=2.0,
                 sequence_tis_ratio_clip_high=5.0, outlier_token_is_threshold_low=out_lo,
                 outlier_token_is_threshold_high=out_hi, geo_mask_low=0.9, geo_mask_high=1.1,
                 product_mask_low=0.5, product_mask_high=2.0)
        cfg = O.make_cfg(policy_loss_type="regular", loss_reduction="token_mean", off_policy_correction=opc)
        x = lp.clone().requires_grad_(True)
        ref, met_ref = O.policy_loss_with_correction(x, olp, adv, cfg, mask, rlp)
        ref.backward()
        y = lp.to(DEV).requires_grad_(True)
        loss, met = skyrl_b200.ppo_policy_loss(y, olp.to(DEV), adv.to(DEV), cfg, loss_mask=mask.to(DEV),
                                               rollout_logprobs=rlp.to(DEV))
        loss.backward()
        assert_close(loss, ref, what="loss with off-policy correction")
        assert_close(y.grad, x.grad, floor=1e-3, what="gradient with off-policy correction")
        for k, v in met_ref.items():
            if k in met and not (np.isnan(v) and np.isnan(met[k])):
                assert met[k] == pytest.approx(v, rel=1e-4, abs=1e-5), k

    run()
