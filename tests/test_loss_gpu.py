"""CUDA policy-loss kernel (through the C ABI) vs the golden reference outputs and the CPU oracle."""
import numpy as np
import pytest
import torch

import skyrl_b200
from skyrl_b200 import losses
from oracle import policy_oracle as O
from conftest import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _t(x):
    return torch.from_numpy(np.asarray(x))


@pytest.mark.parametrize("kind", ["regular", "dual_clip", "gspo"])
@pytest.mark.parametrize("red", ["token_mean", "sequence_mean", "seq_mean_token_sum_norm"])
@pytest.mark.parametrize("mname", ["mask", "nomask"])
def test_registry_loss_matches_reference_golden(golden_loss, kind, red, mname):
    g = golden_loss
    cfg = O.make_cfg(policy_loss_type=kind, loss_reduction=red, max_seq_len=64,
                     eps_clip_high=0.28 if kind == "dual_clip" else 0.2)
    lp = _t(g["lp"]).to(DEV).requires_grad_(True)
    mask = _t(g["mask"]).to(DEV) if mname == "mask" else None
    fn = skyrl_b200.PolicyLossRegistry.get(kind)
    loss, met = fn(lp, _t(g["olp"]).to(DEV), _t(g["adv"]).to(DEV), cfg, loss_mask=mask)
    assert loss.dim() == 0 and isinstance(met["clip_ratio"], float)
    loss.backward()
    key = f"{kind}__{red}__{mname}"
    assert_close(loss, g[key + "__loss"], what=key + " loss")
    assert_close(met["clip_ratio"], g[key + "__clip"], what=key + " clip")
    # gradients at the ratio clamp reach e^20 * |A|: relative tolerance there, floor 1 elsewhere
    assert_close(lp.grad, g[key + "__grad"], what=key + " grad")


@pytest.mark.parametrize("name,kw", [
    ("asm_grpo_kl", dict(policy_loss_type="regular", use_kl_loss=True, use_entropy_loss=False)),
    ("asm_ent", dict(policy_loss_type="dual_clip", use_kl_loss=False, use_entropy_loss=True)),
    ("asm_gspo_all", dict(policy_loss_type="gspo", loss_reduction="sequence_mean", use_kl_loss=True,
                          use_entropy_loss=True, kl_estimator_type="k2")),
])
def test_loss_assembly_matches_reference_golden(golden_loss, name, kw):
    g = golden_loss
    cfg = O.make_cfg(max_seq_len=64, **kw)
    lp = _t(g["lp"]).to(DEV).requires_grad_(True)
    ent = _t(g["ent"]).to(DEV).requires_grad_(True)
    loss, st = losses.policy_update_loss(lp, _t(g["olp"]).to(DEV), _t(g["adv"]).to(DEV), cfg,
                                         loss_mask=_t(g["mask"]).to(DEV), base_log_probs=_t(g["base"]).to(DEV),
                                         entropy=ent)
    loss.backward()
    want = g[name + "__stats"]
    assert_close(st["final_loss"], want[0], what="final")
    assert_close(loss, want[0], what="final tensor")
    assert_close(st["policy_loss"], want[1], what="policy")
    assert_close(st["policy_entropy"], want[2], what="entropy")
    assert_close(st.get("policy_kl", 0.0), want[3], what="kl")
    assert_close(st["loss_metrics/clip_ratio"], want[4], what="clip")
    assert_close(lp.grad, g[name + "__dlp"], what="d_lp")  # KL term contributes no gradient (trap 6)
    assert_close(ent.grad if ent.grad is not None else torch.zeros_like(ent), g[name + "__dent"], what="d_ent")


@pytest.mark.parametrize("B,A", [(1, 1), (3, 7), (5, 4096), (64, 1024), (2, 5000), (7, 4099)])
@pytest.mark.parametrize("kind,red", [("regular", "token_mean"), ("dual_clip", "sequence_mean"),
                                      ("gspo", "sequence_mean"), ("gspo", "token_mean")])
def test_random_shapes_vs_oracle(B, A, kind, red):
    g = torch.Generator().manual_seed(B * 100003 + A)
    lens = torch.randint(1, A + 1, (B,), generator=g)
    mask = (torch.arange(A)[None] < lens[:, None]).float()
    if B > 2:
        mask[1] = 0
    lp = -torch.rand(B, A, generator=g) * 3
    olp = lp + 0.1 * torch.randn(B, A, generator=g)
    adv = torch.randn(B, A, generator=g)
    cfg = O.make_cfg(policy_loss_type=kind, loss_reduction=red, eps_clip_low=0.2, eps_clip_high=0.28, clip_ratio_c=3.0)
    x = lp.clone().requires_grad_(True)
    ref, clip = O.policy_loss(x, olp, adv, cfg, mask)
    ref.backward()
    y = lp.to(DEV).requires_grad_(True)
    loss, met = skyrl_b200.PolicyLossRegistry.get(kind)(y, olp.to(DEV), adv.to(DEV), cfg, loss_mask=mask.to(DEV))
    (loss * 2.5).backward()  # upstream gradient is applied on the device
    assert_close(loss, ref, what="loss")
    assert_close(met["clip_ratio"], clip, what="clip")
    assert_close(y.grad, 2.5 * x.grad, what="grad")


def test_fully_masked_batch_is_finite_zero():  # reference: tests/cpu/algorithms/test_losses.py:242-288
    lp = torch.randn(3, 9, device=DEV, requires_grad=True)
    z = torch.zeros(3, 9, device=DEV)
    for kind, red in (("regular", "token_mean"), ("gspo", "sequence_mean")):
        loss, met = skyrl_b200.PolicyLossRegistry.get(kind)(lp, lp.detach() + 0.1, torch.ones_like(z),
                                                             O.make_cfg(policy_loss_type=kind, loss_reduction=red),
                                                             loss_mask=z)
        assert torch.isfinite(loss) and loss.item() == 0.0 and met["clip_ratio"] == 0.0


def test_integer_mask_and_noncontiguous_inputs():
    g = torch.Generator().manual_seed(3)
    lp = -torch.rand(4, 20, generator=g)
    olp = lp + 0.05 * torch.randn(4, 20, generator=g)
    adv = torch.randn(4, 20, generator=g)
    mask = (torch.rand(4, 20, generator=g) > 0.3).long()
    cfg = O.make_cfg()
    ref, _ = O.policy_loss(lp, olp, adv, cfg, mask.float())
    big = torch.zeros(4, 40, device=DEV)
    big[:, ::2] = lp.to(DEV)
    loss, _ = losses.ppo_policy_loss(big[:, ::2].requires_grad_(True), olp.to(DEV), adv.to(DEV), cfg,
                                     loss_mask=mask.to(DEV))
    assert_close(loss, ref)


def test_disabled_off_policy_correction_is_a_no_op():
    lp = torch.zeros(2, 4, device=DEV, requires_grad=True)
    # correction disabled + rollout logprobs given = the reference's no-op path (off_policy_correction_utils.py:223)
    loss, met = losses.ppo_policy_loss(lp, lp.detach(), lp.detach(), O.make_cfg(), rollout_logprobs=lp.detach() - 1)
    assert loss.item() == 0.0 and list(met) == ["clip_ratio"]


def test_full_size_properties_c2_shape():
    """C2 shape (256 x 4096): checks that do not need the oracle at full size."""
    B, A = 256, 4096
    g = torch.Generator(device=DEV).manual_seed(11)
    lp = -torch.rand(B, A, device=DEV, generator=g)
    adv = torch.randn(B, A, device=DEV, generator=g)
    mask = (torch.rand(B, A, device=DEV, generator=g) > 0.2).float()
    cfg = O.make_cfg()
    x = lp.clone().requires_grad_(True)
    loss, met = losses.ppo_policy_loss(x, lp, adv, cfg, loss_mask=mask)  # on-policy: ratio == 1 everywhere
    loss.backward()
    n = mask.sum()
    assert met["clip_ratio"] == 0.0
    assert_close(loss, -(adv * mask).sum() / n, rel=1e-4)
    assert_close(x.grad, -(adv * mask) / n, rel=1e-5, floor=1e-6)
    # linearity in the advantages
    loss2, _ = losses.ppo_policy_loss(lp.clone().requires_grad_(True), lp, 3 * adv, cfg, loss_mask=mask)
    assert_close(loss2, 3 * loss.detach(), rel=1e-5)


@pytest.mark.parametrize("cname,clip", [("clip02", 0.2), ("noclip", None)])
@pytest.mark.parametrize("mname", ["mask", "nomask"])
def test_critic_loss_matches_reference_golden(golden_loss, cname, clip, mname):
    from types import SimpleNamespace

    g = golden_loss
    v = _t(g["values"]).to(DEV).requires_grad_(True)
    mask = _t(g["mask"]).to(DEV) if mname == "mask" else None
    loss, clipfrac = skyrl_b200.ppo_critic_loss(v, _t(g["old_values"]).to(DEV), _t(g["returns"]).to(DEV),
                                                SimpleNamespace(value_clip=clip), loss_mask=mask)
    loss.backward()
    key = f"critic__{cname}__{mname}"
    assert_close(loss, g[key + "__loss"], what="loss")
    assert (clipfrac is None) == (clip is None)
    if clip is not None:
        assert_close(clipfrac, g[key + "__clipfrac"], what="clipfrac")
    assert_close(v.grad, g[key + "__grad"], what="grad")


_XCFG = dict(sapo=__import__("types").SimpleNamespace(tau_pos=1.0, tau_neg=2.0),
             cispo=__import__("types").SimpleNamespace(cispo_eps_clip_low=0.1, cispo_eps_clip_high=0.3))


@pytest.mark.parametrize("kind,red", [("sapo", "sequence_mean"), ("sapo", "token_mean"), ("cispo", "token_mean"),
                                      ("cispo", "seq_mean_token_sum_norm"), ("cross_entropy", "token_mean"),
                                      ("importance_sampling", "token_mean")])
@pytest.mark.parametrize("mname", ["mask", "nomask"])
def test_extra_registry_losses_match_reference_golden(golden_loss, kind, red, mname):
    g = golden_loss
    cfg = O.make_cfg(policy_loss_type=kind, loss_reduction=red, max_seq_len=64, **_XCFG)
    lp = _t(g["lp"]).to(DEV).requires_grad_(True)
    olp = _t(g["olp_mild"] if kind == "importance_sampling" else g["olp"]).to(DEV)
    mask = _t(g["mask"]).to(DEV) if mname == "mask" else None
    loss, met = skyrl_b200.PolicyLossRegistry.get(kind)(lp, olp, _t(g["adv"]).to(DEV), cfg, loss_mask=mask)
    loss.backward()
    key = f"{kind}__{red}__{mname}"
    assert_close(loss, g[key + "__loss"], what=key + " loss")
    assert_close(list(met.values())[0], g[key + "__metric"], what=key + " metric")
    assert_close(lp.grad, g[key + "__grad"], what=key + " grad")


def test_off_policy_correction_matches_reference_golden(golden_loss):
    import json
    from types import SimpleNamespace

    g = golden_loss
    keys = sorted(k[:-6] for k in g if k.startswith("opc__") and k.endswith("__loss"))
    assert len(keys) == 8
    for key in keys:
        okw, want_met = json.loads(str(g[key + "__cfg"])), json.loads(str(g[key + "__metrics"]))
        kind = key.split("__")[-1]
        cfg = O.make_cfg(policy_loss_type=kind, loss_reduction="token_mean" if kind == "regular" else "sequence_mean",
                         eps_clip_high=0.28 if kind == "dual_clip" else 0.2, max_seq_len=64,
                         off_policy_correction=SimpleNamespace(**okw))
        lp = _t(g["lp"]).to(DEV).requires_grad_(True)
        loss, met = skyrl_b200.PolicyLossRegistry.get(kind)(lp, _t(g["olp"]).to(DEV), _t(g["adv"]).to(DEV), cfg,
                                                             loss_mask=_t(g["mask"]).to(DEV),
                                                             rollout_logprobs=_t(g["rlp"]).to(DEV))
        loss.backward()
        assert_close(loss, g[key + "__loss"], what=key + " loss")
        assert_close(lp.grad, g[key + "__grad"], what=key + " grad")
        assert sorted(met) == sorted(want_met), key  # same metric names as the reference
        for k in want_met:
            assert met[k] == pytest.approx(want_met[k], rel=1e-5, abs=1e-6), (key, k)


def test_off_policy_correction_known_answers():  # reference: tests/cpu/algorithms/test_off_policy_correction.py
    import opc_kat

    for name, olp, rlp, mask, kw, want_tis, want_keep, want_met in opc_kat.CASES:
        olp_t, rlp_t, m_t = (torch.tensor(x, device=DEV) for x in (olp, rlp, mask))
        tis, met, new_mask = losses.compute_off_policy_correction(olp_t, rlp_t, m_t, opc_kat.cfg(**kw))
        want_mask = torch.tensor(mask) * torch.tensor(want_keep).unsqueeze(-1)
        assert torch.equal(new_mask.cpu(), want_mask), name
        if want_tis is not None:
            assert torch.allclose(tis.cpu(), torch.tensor(want_tis), rtol=1e-3, atol=1e-4), name
        for k, v in want_met.items():
            assert abs(met[k] - v) < 0.01, (name, k, met[k], v)
    with pytest.raises(ValueError, match="Unknown tis_ratio_type"):
        losses.compute_off_policy_correction(olp_t, rlp_t, m_t, opc_kat.cfg(tis_ratio_type="bogus"))
    with pytest.raises(ValueError, match="Unknown sequence_mask_metric"):
        losses.compute_off_policy_correction(olp_t, rlp_t, m_t, opc_kat.cfg(sequence_mask_metric="bogus"))


# ------------------------------------------------------------------------------------ clip_cov / kl_cov
_COV_CFG = dict(clip_cov=__import__("types").SimpleNamespace(clip_ratio=0.1, clip_cov_lb=0.05, clip_cov_ub=5.0),
                kl_cov=__import__("types").SimpleNamespace(kl_cov_frac=0.2, ppo_kl_coef=1.0))


@pytest.mark.parametrize("red", ["token_mean", "seq_mean_token_sum_norm"])
def test_kl_cov_matches_reference_golden(golden_loss, red):
    g = golden_loss
    cfg = O.make_cfg(policy_loss_type="kl_cov", loss_reduction=red, max_seq_len=64, **_COV_CFG)
    lp = _t(g["lp"]).to(DEV).requires_grad_(True)
    loss, met = skyrl_b200.PolicyLossRegistry.get("kl_cov")(lp, _t(g["olp_mild"]).to(DEV), _t(g["adv"]).to(DEV), cfg,
                                                            loss_mask=_t(g["mask"]).to(DEV))
    loss.backward()
    key = f"kl_cov__{red}__mask"
    assert met == {"clip_ratio": 0.0}
    assert_close(loss, g[key + "__loss"], what=key + " loss")
    assert_close(lp.grad, g[key + "__grad"], what=key + " grad")


def _cov_batch(B, A, seed):
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(1, A + 1, (B,), generator=g)
    mask = (torch.arange(A)[None] < lens[:, None]).float()
    mask[B // 2] = 0.0
    lp = -torch.rand(B, A, generator=g) * 3
    olp = lp + 0.2 * torch.randn(B, A, generator=g)
    adv = torch.randn(B, A, generator=g)
    return lp, olp, adv, mask


@pytest.mark.parametrize("shape", [(7, 33), (64, 1000), (16, 8192)])
def test_kl_cov_selection_is_the_oracle_topk(shape):
    """The radix select picks exactly the oracle's torch.topk set (means differ by an ulp at most: allow 2 swaps)."""
    B, A = shape
    lp, olp, adv, mask = _cov_batch(B, A, 5 + A)
    cfg = O.make_cfg(policy_loss_type="kl_cov", loss_reduction="token_mean", **_COV_CFG)
    want = O.cov_selection("kl_cov", lp, olp, adv, cfg, mask)
    sel, counts = losses.cov_token_selection("kl_cov", lp.to(DEV), olp.to(DEV), adv.to(DEV), cfg, mask.to(DEV))
    k = max(1, int(int((mask > 0).sum()) * 0.2))
    assert counts.tolist() == [float(k), float((mask > 0).sum())]
    assert int(sel.sum().item()) == k  # exactly k, never k +- ties
    assert int((sel.cpu() != want).sum()) <= 2
    x = lp.clone().to(DEV).requires_grad_(True)
    loss, _ = losses.compute_policy_loss_kl_cov(x, olp.to(DEV), adv.to(DEV), cfg, loss_mask=mask.to(DEV))
    loss.backward()
    xr = lp.clone().requires_grad_(True)
    ref, _, _ = O.cov_policy_loss("kl_cov", xr, olp, adv, cfg, mask, sel=sel.cpu())
    ref.backward()
    assert_close(loss, ref, what="kl_cov loss")
    assert_close(x.grad, xr.grad, what="kl_cov grad")


def test_kl_cov_ties_and_empty_mask():
    B, A = 4, 100
    lp, olp, adv, mask = _cov_batch(B, A, 3)
    cfg = O.make_cfg(policy_loss_type="kl_cov", loss_reduction="token_mean", **_COV_CFG)
    const_adv = torch.ones(B, A)  # adv - mean = 0 -> every cov ties at 0: ties are broken by index, count stays exact
    sel, counts = losses.cov_token_selection("kl_cov", lp.to(DEV), olp.to(DEV), const_adv.to(DEV), cfg, mask.to(DEV))
    k = max(1, int(int((mask > 0).sum()) * 0.2))
    flat_valid = torch.nonzero((mask > 0).reshape(-1)).squeeze(-1)
    assert int(sel.sum().item()) == k
    assert torch.equal(torch.nonzero(sel.cpu().reshape(-1)).squeeze(-1), flat_valid[:k])
    sel0, counts0 = losses.cov_token_selection("kl_cov", lp.to(DEV), olp.to(DEV), adv.to(DEV), cfg,
                                               torch.zeros(B, A, device=DEV))
    assert sel0.sum().item() == 0 and counts0.tolist() == [0.0, 0.0]
    loss, _ = losses.compute_policy_loss_kl_cov(lp.to(DEV), olp.to(DEV), adv.to(DEV), cfg,
                                                loss_mask=torch.zeros(B, A, device=DEV))
    assert loss.item() == 0.0


@pytest.mark.parametrize("shape,red", [((12, 40), "token_mean"), ((64, 1000), "sequence_mean")])
def test_clip_cov_selection_properties_and_loss(shape, red):
    """clip_cov draws a random subset (the reference uses torch.randperm, ppo_utils.py:806), so parity is: the subset
    lies inside the oracle's candidate set, has exactly min(clip_num, #candidates) tokens, and loss / gradient /
    clip_ratio equal the oracle's for that same subset."""
    B, A = shape
    lp, olp, adv, mask = _cov_batch(B, A, 17 + A)
    cfg = O.make_cfg(policy_loss_type="clip_cov", loss_reduction=red, **_COV_CFG)
    cand = O.cov_selection("clip_cov", lp, olp, adv, cfg, mask, perm_fn=lambda n: torch.arange(n))
    c = cfg.clip_cov  # recompute the full candidate set (the oracle call above returns only the first clip_num)
    full = O.cov_selection("clip_cov", lp, olp, adv, O.make_cfg(
        policy_loss_type="clip_cov", clip_cov=__import__("types").SimpleNamespace(
            clip_ratio=1e9, clip_cov_lb=c.clip_cov_lb, clip_cov_ub=c.clip_cov_ub)), mask,
        perm_fn=lambda n: torch.arange(n))
    n_cand = int(full.sum())
    k = min(max(int(c.clip_ratio * mask.sum().item()), 1), n_cand)
    assert int(cand.sum()) == k and n_cand > k
    args = (lp.to(DEV), olp.to(DEV), adv.to(DEV), cfg, mask.to(DEV))
    sel, counts = losses.cov_token_selection("clip_cov", *args, seed=1234)
    assert counts.tolist() == [float(k), float(n_cand)]
    assert int(sel.sum().item()) == k and torch.all(full[sel.cpu() > 0] == 1)
    again, _ = losses.cov_token_selection("clip_cov", *args, seed=1234)
    other, _ = losses.cov_token_selection("clip_cov", *args, seed=99)
    assert torch.equal(sel, again) and not torch.equal(sel, other)
    # the registry function: seeded through torch's global generator like the reference
    torch.manual_seed(5)
    x = lp.clone().to(DEV).requires_grad_(True)
    loss, met = skyrl_b200.PolicyLossRegistry.get("clip_cov")(x, olp.to(DEV), adv.to(DEV), cfg, loss_mask=mask.to(DEV))
    loss.backward()
    torch.manual_seed(5)
    seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    sel5, _ = losses.cov_token_selection("clip_cov", *args, seed=seed)
    xr = lp.clone().requires_grad_(True)
    ref, ref_met, _ = O.cov_policy_loss("clip_cov", xr, olp, adv, cfg, mask, sel=sel5.cpu())
    ref.backward()
    assert_close(loss, ref, what="clip_cov loss")
    assert_close(met["clip_ratio"], ref_met["clip_ratio"], what="clip_cov clip_ratio")
    assert_close(x.grad, xr.grad, what="clip_cov grad")


def test_clip_cov_draw_is_uniform_over_candidates():
    """Every candidate is dropped with probability k / #candidates: frequency over 200 seeds within 5 sigma."""
    B, A = 8, 64
    lp, olp, adv, mask = _cov_batch(B, A, 2)
    cfg = O.make_cfg(policy_loss_type="clip_cov", loss_reduction="token_mean", **_COV_CFG)
    args = (lp.to(DEV), olp.to(DEV), adv.to(DEV), cfg, mask.to(DEV))
    total = torch.zeros(B, A, device=DEV)
    n = 200
    for s in range(n):
        sel, counts = losses.cov_token_selection("clip_cov", *args, seed=s)
        total += sel
    k, n_cand = counts.tolist()
    p = k / n_cand
    freq = total[total > 0] / n
    assert int((total > 0).sum()) > 0.9 * n_cand  # nearly every candidate shows up at some point
    assert (freq - p).abs().max().item() < 5 * (p * (1 - p) / n) ** 0.5 + 1e-6
