"""CUDA advantage kernels (GRPO / RLOO / GAE / KL) vs the golden reference outputs and the CPU oracle."""
import numpy as np
import pytest
import torch

import skyrl_b200
from skyrl_b200 import advantages
from oracle import policy_oracle as O
from conftest import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _t(x):
    return torch.from_numpy(np.asarray(x))


def test_grpo_rloo_match_reference_golden(golden_adv):
    g = golden_adv
    r, dense = _t(g["rewards"]), _t(g["dense_rewards"])
    im = _t(g["mask_i64"])
    fm = im.float()
    uids = g["uids"]
    est = skyrl_b200.AdvantageEstimatorRegistry.get
    # CPU tensors in (as the reference's driver passes them), int64 mask consumed natively
    adv, ret = est("grpo")(token_level_rewards=r, response_mask=im, index=uids, values=None, grpo_norm_by_std=True,
                           gamma=1.0, lambd=1.0, config=None)
    assert adv is ret and adv.device.type == "cpu" and adv.dtype == torch.float32
    assert_close(adv, g["grpo_std"], what="grpo_std")
    assert torch.equal(adv == 0, _t(g["grpo_std"]) == 0), "mask zeros must be bit-exact"
    adv, _ = est("grpo")(token_level_rewards=r.to(DEV), response_mask=fm.to(DEV), index=list(uids),
                         grpo_norm_by_std=False)
    assert adv.is_cuda
    assert_close(adv, g["grpo_nostd"], what="grpo_nostd")
    assert_close(est("grpo")(token_level_rewards=dense, response_mask=fm, index=uids)[0], g["grpo_dense"])
    assert_close(est("rloo")(token_level_rewards=r, response_mask=fm, index=uids)[0], g["rloo"], what="rloo")
    assert_close(est("rloo")(token_level_rewards=dense.to(DEV), response_mask=im.to(DEV), index=uids)[0],
                 g["rloo_dense"], what="rloo dense")


def test_gae_matches_reference_golden(golden_adv):
    g = golden_adv
    dense, v, fm = _t(g["dense_rewards"]), _t(g["values"]), _t(g["mask_i64"]).float()
    for name, (gm, lm) in {"gae_1_095": (1.0, 0.95), "gae_099_1": (0.99, 1.0), "gae_1_1": (1.0, 1.0)}.items():
        adv, ret = skyrl_b200.compute_advantages_and_returns(dense.to(DEV), fm.to(DEV), None, "gae", None,
                                                             values=v.to(DEV), gamma=gm, lambd=lm)
        assert_close(adv, g[name + "_adv"], what=name + " adv")
        assert_close(ret, g[name + "_ret"], what=name + " ret")


def test_gae_known_answers():  # reference: tests/cpu/utils/test_ppo_utils.py:175-242
    r, v = torch.tensor([[1.0, 2.0, 3.0]]), torch.tensor([[0.5, 1.0, 1.5]])
    adv, ret = advantages.compute_gae_advantage_return(r, v, torch.tensor([[1.0, 0.0, 1.0]]), 1.0, 1.0)
    assert torch.allclose(ret, torch.tensor([[6.0, 5.0, 3.0]]), atol=1e-5)
    assert torch.allclose(adv, torch.tensor([[0.7071, 0.1768, -0.7071]]), atol=1e-4)  # no reset on the mask
    _, ret = advantages.compute_gae_advantage_return(r, v, torch.ones(1, 3), 0.5, 1.0)
    assert torch.allclose(ret, torch.tensor([[2.75, 3.5, 3.0]]), atol=1e-5)
    _, ret = advantages.compute_gae_advantage_return(r, v, torch.ones(1, 3), 1.0, 0.5)
    assert torch.allclose(ret, torch.tensor([[3.625, 4.25, 3.0]]), atol=1e-5)


def test_gae_mask_errors_like_masked_var():
    r = torch.zeros(2, 3)
    with pytest.raises(ValueError, match="At least one element"):
        advantages.compute_gae_advantage_return(r, r, torch.zeros(2, 3), 1.0, 1.0)
    m = torch.zeros(2, 3)
    m[0, 0] = 1
    with pytest.raises(ValueError, match="sum of the mask is one"):
        advantages.compute_gae_advantage_return(r, r, m, 1.0, 1.0)


@pytest.mark.parametrize("B,A", [(1, 1), (3, 17), (4, 4096), (5, 4097), (2, 9000), (16, 8192)])
@pytest.mark.parametrize("gamma,lambd", [(1.0, 0.95), (1.0, 1.0), (0.99, 0.9)])
def test_gae_random_vs_oracle(B, A, gamma, lambd):
    if B * A < 2:
        pytest.skip("masked_var needs two elements")
    g = torch.Generator().manual_seed(A * 7 + B)
    lens = torch.randint(max(1, A // 2), A + 1, (B,), generator=g)
    mask = (torch.arange(A)[None] < lens[:, None]).float()
    r = torch.zeros(B, A)
    r[torch.arange(B), lens - 1] = torch.randn(B, generator=g)
    r = r - 0.01 * torch.rand(B, A, generator=g) * mask
    v = torch.randn(B, A, generator=g)
    a_ref, r_ref = O.gae_advantage_return(r, v, mask, gamma, lambd)
    a, ret = advantages.compute_gae_advantage_return(r.to(DEV), v.to(DEV), mask.to(DEV), gamma, lambd)
    assert_close(ret, r_ref, what="returns")
    assert_close(a, a_ref, what="advantages")


@pytest.mark.parametrize("B,A,G", [(8, 5, 4), (64, 1024, 8), (48, 1000, 16), (6, 4100, 1)])
def test_group_random_vs_oracle(B, A, G):
    g = torch.Generator().manual_seed(B + A)
    uids = np.array([f"uid-{i // G}" for i in range(B)])
    uids = uids[torch.randperm(B, generator=g).numpy()]  # non-contiguous groups
    lens = torch.randint(1, A + 1, (B,), generator=g)
    mask = (torch.arange(A)[None] < lens[:, None]).long()
    r = torch.zeros(B, A)
    r[torch.arange(B), lens - 1] = (torch.rand(B, generator=g) < 0.3).float()
    for mode, fn, kw in (("grpo", O.grpo_advantage, {}), ("grpo", O.grpo_advantage, {"norm_by_std": False}),
                         ("rloo", O.rloo_advantage, {})):
        want, _ = fn(r, mask, uids, **kw)
        got, _ = skyrl_b200.AdvantageEstimatorRegistry.get(mode)(
            token_level_rewards=r.to(DEV), response_mask=mask.to(DEV), index=uids,
            grpo_norm_by_std=kw.get("norm_by_std", True))
        assert_close(got, want, what=f"{mode} {kw}")
        assert torch.equal(got.cpu() == 0, want == 0)


@pytest.mark.parametrize("B,A,G", [(64, 1024, 8), (48, 1000, 16), (6, 4100, 1), (16, 8192, 8), (8, 32768, 8),
                                   (32, 12000, 4), (10, 16384, 2), (32, 16384, 16), (21, 8, 7)])
@pytest.mark.parametrize("int_mask", [False, True])
def test_group_cluster_form_vs_general_form_and_oracle(B, A, G, int_mask):
    """Uniform prompt groups take the one-launch thread-block-cluster kernel; the same batch through the general
    two-kernel form and the oracle must agree (zero pattern bit-exact)."""
    g = torch.Generator().manual_seed(B * 31 + A + G)
    uids = np.array([f"uid-{i // G}" for i in range(B)])[torch.randperm(B, generator=g).numpy()]
    lens = torch.randint(1, A + 1, (B,), generator=g)
    mask = (torch.arange(A)[None] < lens[:, None])
    mask = mask.long() if int_mask else mask.float()
    r = 0.01 * torch.randn(B, A, generator=g)
    r[torch.arange(B), lens - 1] += (torch.rand(B, generator=g) < 0.5).float()
    for mode, fn, kw in ((0, O.grpo_advantage, {}), (1, O.grpo_advantage, {"norm_by_std": False}),
                         (2, O.rloo_advantage, {})):
        want, _ = fn(r, mask, uids, **kw)
        one = advantages._group_advantage(r.to(DEV), mask.to(DEV), uids, mode, 1e-6 if mode < 2 else 0.0, one_launch=True)
        two = advantages._group_advantage(r.to(DEV), mask.to(DEV), uids, mode, 1e-6 if mode < 2 else 0.0, one_launch=False)
        assert_close(one, want, what=f"cluster form, mode {mode}")
        assert_close(one, two, what=f"cluster vs general form, mode {mode}")
        assert torch.equal(one.cpu() == 0, want == 0)


def test_group_cluster_form_refuses_ragged_groups():
    r, m = torch.zeros(3, 8, device=DEV), torch.ones(3, 8, device=DEV)
    with pytest.raises(ValueError, match="uniform groups"):
        advantages._group_advantage(r, m, ["a", "a", "b"], 0, 1e-6, one_launch=True)
    with pytest.raises(ValueError, match="uniform groups"):  # A not a multiple of 4
        advantages._group_advantage(r[:, :5].contiguous(), m[:, :5].contiguous(), ["a", "b", "c"], 0, 1e-6, one_launch=True)


def test_kl_estimators_and_reward_penalty_match_reference_golden(golden_adv):
    g = golden_adv
    lp, base, fm, r = _t(g["lp"]), _t(g["base"]), _t(g["mask_i64"]).float(), _t(g["rewards"])
    for kind in ("k1", "abs", "k2", "k3"):
        assert_close(skyrl_b200.compute_approx_kl(lp.to(DEV), base.to(DEV), fm.to(DEV), kind), g["kl_" + kind])
        assert_close(skyrl_b200.compute_approx_kl(lp, base, None, kind), g["kl_nomask_" + kind])
        out, met = skyrl_b200.apply_reward_kl_penalty(r.to(DEV), lp.to(DEV), base.to(DEV), fm.to(DEV), kind, 0.05)
        assert_close(out, g["klrew_" + kind], what="klrew " + kind)
        assert_close([met["avg_kl"], met["avg_kl_max"]], g["klrew_metrics_" + kind], what="klrew metrics")
    out, _ = skyrl_b200.apply_reward_kl_penalty(r.to(DEV), lp.to(DEV), base.to(DEV), fm.to(DEV), "k3", -1.0)
    assert torch.equal(out.cpu(), r), "negative coefficient is clamped to 0 (max(0, coef))"


def test_kl_in_reward_known_answer():  # reference: tests/cpu/test_trainer.py:66-133 (k3): 0.1249 / 0.3143
    probs = torch.tensor([[0.1, 0.2, 0.3, 0.2, 0.2], [0.25, 0.25, 0.25, 0.15, 0.10]])
    base = torch.tensor([[0.15, 0.15, 0.25, 0.25, 0.2], [0.2, 0.3, 0.2, 0.2, 0.10]])
    lp, bp = probs.log(), base.log()
    mask = torch.ones(2, 5)
    want_out, want = O.reward_kl_penalty(torch.zeros(2, 5), lp, bp, mask, "k3", 0.1)
    out, met = skyrl_b200.apply_reward_kl_penalty(torch.zeros(2, 5), lp, bp, mask, "k3", 0.1)
    assert_close(out, want_out)
    assert met["avg_kl"] == pytest.approx(want["avg_kl"], rel=1e-5)
    assert met["avg_kl_max"] == pytest.approx(want["avg_kl_max"], rel=1e-5)


def test_full_size_properties_c3_shape():
    """C3 per-GPU shard (64 x 8192 here x8 = 512 x 8192): size-independent checks of GAE + GRPO."""
    B, A = 512, 8192
    g = torch.Generator(device=DEV).manual_seed(5)
    v = torch.randn(B, A, device=DEV, generator=g)
    r = torch.zeros(B, A, device=DEV)
    r[:, -1] = 1.0
    mask = torch.ones(B, A, device=DEV)
    adv, ret = advantages.compute_gae_advantage_return(r, v, mask, 1.0, 1.0)
    # gamma = lambda = 1: returns_t = sum_{s>=t} r_s = 1 (telescoping), independent of the values.  delta_t is
    # formed in fp32 exactly as the reference forms it, so 8192 rounded deltas leave ~1e-5 of noise around 1.
    assert_close(ret, torch.ones(B, A), rel=1e-4)
    # whitening: zero mean, unit (Bessel) variance over the mask
    assert abs(adv.mean().item()) < 1e-4 and abs(adv.var().item() - 1.0) < 1e-3
    uids = [i // 8 for i in range(B)]
    a, _ = advantages.compute_grpo_outcome_advantage(r, mask, uids)  # every group has equal scores -> 0
    assert torch.count_nonzero(a) == 0


def test_reinforce_pp_and_normalize_match_reference_golden(golden_adv):
    g = golden_adv
    dense, fm = _t(g["dense_rewards"]), _t(g["mask_i64"]).float()
    est = skyrl_b200.AdvantageEstimatorRegistry.get("reinforce++")
    for gm in (1.0, 0.9):
        adv, ret = est(token_level_rewards=dense.to(DEV), response_mask=fm.to(DEV), index=None, values=None,
                       gamma=gm, lambd=1.0, config=None)
        assert_close(ret, g[f"rpp_{gm}_ret"], what="rpp ret")
        assert_close(adv, g[f"rpp_{gm}_adv"], what="rpp adv")
    adv, ret = est(token_level_rewards=dense, response_mask=_t(g["rpp_hole_mask"]), gamma=0.95)
    assert_close(ret, g["rpp_hole_ret"], what="rpp hole ret")  # running return resets inside the hole
    assert_close(adv, g["rpp_hole_adv"], what="rpp hole adv")
    assert torch.equal(adv == 0, _t(g["rpp_hole_adv"]) == 0)
    assert_close(skyrl_b200.normalize_advantages(_t(g["norm_in"]).to(DEV), fm.to(DEV)), g["norm_out"])
    assert_close(skyrl_b200.normalize_advantages(_t(g["norm_in"]), _t(g["mask_i64"])), g["norm_out"])


def test_reinforce_pp_known_answers_and_long_rows():  # reference: tests/cpu/utils/test_ppo_utils.py:65-100
    r = torch.tensor([[1.0, 2.0, 3.0]])
    _, ret = advantages.compute_reinforce_plus_plus_outcome_advantage(r, torch.tensor([[1.0, 1.0, 0.0]]), 1.0)
    assert torch.allclose(ret, torch.tensor([[3.0, 2.0, 3.0]]), atol=1e-5)
    _, ret = advantages.compute_reinforce_plus_plus_outcome_advantage(r, torch.ones(1, 3), 0.5)
    assert torch.allclose(ret, torch.tensor([[2.75, 3.5, 3.0]]), atol=1e-5)
    g = torch.Generator().manual_seed(9)
    B, A = 6, 9001  # several 4096-position tiles with carries across them, holes in the mask
    mask = (torch.rand(B, A, generator=g) > 0.01).float()
    rew = torch.randn(B, A, generator=g) * 0.1
    a_ref, r_ref = O.reinforce_pp_advantage(rew, mask, 0.999)
    a, ret = advantages.compute_reinforce_plus_plus_outcome_advantage(rew.to(DEV), mask.to(DEV), 0.999)
    assert_close(ret, r_ref, what="ret")
    assert_close(a, a_ref, what="adv")


def test_step_wise_trajectory_broadcast_and_metrics():
    """trainer.py:777-808: the estimator runs on last-step rows only, every step row of a trajectory gets them."""
    from types import SimpleNamespace

    from skyrl_b200 import trainer_ops

    g = torch.Generator().manual_seed(12)
    steps = [3, 1, 2, 4, 2, 1, 3, 2]  # trajectories of a multi-turn rollout, rows = steps
    Bt, A = sum(steps), 40
    is_last = torch.zeros(Bt, dtype=torch.bool)
    is_last[torch.tensor(steps).cumsum(0) - 1] = True
    traj_uid = [f"prompt-{t // 2}" for t in range(len(steps))]  # two trajectories per prompt
    uids = [u for u, n in zip(traj_uid, steps) for _ in range(n)]
    lens = torch.randint(5, A + 1, (Bt,), generator=g)
    mask = (torch.arange(A)[None] < lens[:, None]).long()
    rewards = torch.zeros(Bt, A)
    rewards[is_last, lens[is_last] - 1] = torch.rand(int(is_last.sum()), generator=g)
    data = {"rewards": rewards.to(DEV), "response_mask": mask.to(DEV), "values": None, "is_last_step": is_last.to(DEV)}
    cfg = SimpleNamespace(advantage_estimator="grpo", gamma=1.0, lambd=1.0, grpo_norm_by_std=True)
    data, met = trainer_ops.compute_advantages_and_returns_batch(data, uids, cfg, step_wise_trajectories=True)
    adv_last, _ = O.grpo_advantage(rewards[is_last], mask[is_last], np.array(uids)[is_last.numpy()])
    traj_ids = torch.cat([torch.tensor([False]), is_last[:-1]]).int().cumsum(0)
    want = adv_last[traj_ids]
    assert_close(data["advantages"], want)
    assert data["returns"] is data["advantages"]
    valid = torch.masked_select(want, mask.bool())
    assert met["avg_advantages"] == pytest.approx(valid.mean().item(), rel=1e-5, abs=1e-6)
    assert met["avg_advantages_abs"] == pytest.approx(valid.abs().mean().item(), rel=1e-5, abs=1e-6)
    assert met["avg_final_rewards"] == pytest.approx(rewards.sum(-1)[is_last].mean().item(), rel=1e-5)
    # plain (non step-wise) path + KL controller plumbing
    ctrl = trainer_ops.AdaptiveKLController(0.05, target=0.1, horizon=100)
    d2 = {"rewards": rewards.to(DEV), "loss_mask": mask.float().to(DEV),
          "action_log_probs": -torch.rand(Bt, A, generator=g).to(DEV),
          "base_action_log_probs": -torch.rand(Bt, A, generator=g).to(DEV)}
    want_r, want_m = O.reward_kl_penalty(rewards, d2["action_log_probs"].cpu(), d2["base_action_log_probs"].cpu(),
                                         mask.float(), "k3", 0.05)
    d2, m2 = trainer_ops.apply_reward_kl_penalty_batch(d2, SimpleNamespace(kl_estimator_type="k3", kl_loss_coef=1e-3), ctrl)
    assert_close(d2["rewards"], want_r)
    assert m2["avg_kl"] == pytest.approx(want_m["avg_kl"], rel=1e-5)
    assert ctrl.value != 0.05  # the controller moved
