"""BASELINE configs 3, 4 and 5 at their real head / sequence shapes (a per-GPU slice of the batch dimension), through
the public API and checked with tests/step_checker.py: the status dict and the loss gradients against the oracle on
the step's own forward outputs, sampled token rows against the fp32 oracle head (forward and d_hidden), and the
size-independent identities over the whole d_weight / d_hidden.  Config 2 at full size lives in
test_policy_update_gpu.py::test_full_size_config2_sampled_rows_and_checksum_identities.

(The file name sorts last on purpose: these are the largest allocations of the suite - up to 27 GB - and run after
everything else.)
"""
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

import skyrl_b200
from oracle import policy_oracle as O
from conftest import assert_close
from step_checker import check_step

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _need_free(gb):
    torch.cuda.empty_cache()  # blocks cached from earlier tests of this process do not count as used
    free, _ = torch.cuda.mem_get_info()
    if free < gb * 1e9:
        pytest.skip(f"needs ~{gb} GB of free HBM")


def _head(N, H, V, gen):
    hidden = torch.empty((N, H), dtype=torch.bfloat16, device=DEV)
    for r0 in range(0, N, 65536):
        hidden[r0:r0 + 65536] = torch.randn((min(65536, N - r0), H), generator=gen, device=DEV)
    weight = torch.empty((V, H), dtype=torch.bfloat16, device=DEV)
    for v0 in range(0, V, 16384):
        n = min(16384, V - v0)
        weight[v0:v0 + n] = torch.randn((n, H), generator=gen, device=DEV) * (1.5 / H ** 0.5)
    labels = torch.randint(0, V, (N,), generator=gen, device=DEV)
    return hidden, weight, labels


def _current_log_probs(hidden, weight, labels, shape, cfg, T, **kw):
    """A loss-only pass: the forward does not depend on the loss inputs, so this yields the current log-probs."""
    z = torch.zeros(shape, device=DEV)
    return skyrl_b200.policy_update_step(hidden, weight, labels, z, z, cfg, need_d_hidden=False, need_d_weight=False,
                                         temperature=T, **kw).log_probs


def test_config3_ppo_gae_kl_in_reward_shape():
    """Config 3 (PPO: GAE gamma=1 lambda=0.95 + KL-in-reward, Qwen2.5-7B head, 8192-token responses), 32 sequences
    (half of one GPU's 64-sequence shard): KL-in-reward and GAE against the oracle on the whole slice, then the step."""
    _need_free(40)
    B, A, H, V, T = 32, 8192, 3584, 152064, 1.0
    gen = torch.Generator(device=DEV).manual_seed(3)
    hidden, weight, labels = _head(B * A, H, V, gen)
    lens = torch.randint(A // 4, A + 1, (B,), generator=gen, device=DEV)
    mask = (torch.arange(A, device=DEV)[None] < lens[:, None]).float()
    cfg = O.make_cfg(policy_loss_type="regular", loss_reduction="token_mean", use_kl_loss=True)
    lp0 = _current_log_probs(hidden, weight, labels, (B, A), cfg, T, loss_mask=mask)
    olp = lp0 + 0.05 * torch.randn((B, A), generator=gen, device=DEV)
    base = lp0 + 0.1 * torch.randn((B, A), generator=gen, device=DEV)
    rewards = torch.zeros(B, A, device=DEV)
    rewards[torch.arange(B, device=DEV), lens - 1] = (torch.rand(B, generator=gen, device=DEV) < 0.3).float()
    values = torch.randn((B, A), generator=gen, device=DEV)  # including the padding positions (SURVEY trap 1)

    r_kl, met = skyrl_b200.apply_reward_kl_penalty(rewards, lp0, base, mask, "k3", 1e-3)
    want_r, want_met = O.reward_kl_penalty(rewards.cpu(), lp0.cpu(), base.cpu(), mask.cpu(), "k3", 1e-3)
    assert_close(r_kl, want_r, what="KL-in-reward")
    assert met["avg_kl"] == pytest.approx(want_met["avg_kl"], rel=1e-4)
    assert met["avg_kl_max"] == pytest.approx(want_met["avg_kl_max"], rel=1e-4)
    adv, ret = skyrl_b200.compute_gae_advantage_return(r_kl, values, mask, 1.0, 0.95)
    want_adv, want_ret = O.gae_advantage_return(r_kl.cpu(), values.cpu(), mask.cpu(), 1.0, 0.95)
    assert_close(ret, want_ret, what="GAE returns")
    assert_close(adv, want_adv, what="GAE advantages")

    out = skyrl_b200.policy_update_step(hidden, weight, labels, olp, adv, cfg, loss_mask=mask, base_log_probs=base,
                                        temperature=T)
    assert torch.allclose(out.log_probs, lp0, atol=2e-5, rtol=0)
    print("config 3:", check_step(out, hidden, weight, labels, olp, adv, cfg, mask, base, T))


@pytest.mark.parametrize("kw", [
    dict(policy_loss_type="gspo", loss_reduction="sequence_mean", eps_clip_low=3e-4, eps_clip_high=4e-4),
    dict(policy_loss_type="gspo", loss_reduction="sequence_mean", eps_clip_low=0.2, eps_clip_high=0.2),
    dict(policy_loss_type="dual_clip", loss_reduction="token_mean", eps_clip_low=0.2, eps_clip_high=0.28,
         clip_ratio_c=3.0),
], ids=["gspo-3e-4", "gspo-0.2", "dual_clip"])
def test_config4_long_horizon_32b_head_shape(kw):
    """Config 4 (dual-clip / GSPO, Qwen3-32B-shaped head H=5120 V=151936, 32K-token trajectories), 2 sequences.
    GSPO chunks are whole sequences (one 32768-token chunk each, a 10 GB logits tile store per parity)."""
    _need_free(60)
    B, A, H, V, T = 2, 32768, 5120, 151936, 1.0
    gen = torch.Generator(device=DEV).manual_seed(4)
    hidden, weight, labels = _head(B * A, H, V, gen)
    lens = torch.randint(24 * 1024, A + 1, (B,), generator=gen, device=DEV)
    mask = (torch.arange(A, device=DEV)[None] < lens[:, None]).float()
    mask[0, 5000:7000] = 0  # tool-call tokens inside the trajectory
    cfg = O.make_cfg(use_kl_loss=False, **kw)
    lp0 = _current_log_probs(hidden, weight, labels, (B, A), cfg, T, loss_mask=mask)
    if kw["policy_loss_type"] == "gspo":
        # sequence log-ratio = masked mean_t(lp - olp) = off_b + O(1e-5): sequence 0 stays inside the clip range (it
        # carries the gradient), sequence 1 leaves it on the side its positive advantage clips
        eps = kw["eps_clip_high"]
        off = torch.tensor([[eps / 4], [3 * eps]], device=DEV)
        olp = lp0 - off + 0.002 * torch.randn((B, A), generator=gen, device=DEV)
        adv = torch.tensor([[-0.7], [1.3]], device=DEV).expand(B, A).contiguous()
    else:
        olp = lp0 + 0.5 * torch.randn((B, A), generator=gen, device=DEV)  # both clips and the dual clip are hit
        adv = torch.randn((B, A), generator=gen, device=DEV)
    out = skyrl_b200.policy_update_step(hidden, weight, labels, olp, adv, cfg, loss_mask=mask, temperature=T)
    assert torch.allclose(out.log_probs, lp0, atol=2e-5, rtol=0)
    assert 0.0 < out.status["loss_metrics/clip_ratio"] < 1.0  # the clip is exercised, and not everywhere
    print("config 4:", kw["policy_loss_type"], check_step(out, hidden, weight, labels, olp, adv, cfg, mask, None, T))


def test_config5_rloo_entropy_bonus_ragged_packed_shape():
    """Config 5 (RLOO, group 16, entropy bonus, ragged response lengths 128-16K, Qwen2.5-7B head): 32 sequences = 2
    groups.  Advantages on the padded [B, A] batch as the trainer calls them, then the step on the PACKED layout
    (no padding row reaches the GEMMs) with the entropy gradient switched on."""
    _need_free(40)
    B, A, H, V, G, T = 32, 16384, 3584, 152064, 16, 1.0
    g = torch.Generator().manual_seed(5)
    lens = torch.randint(128, A + 1, (B,), generator=g)
    lens[0], lens[1] = 128, A  # both ends of the range
    keep = torch.arange(A)[None] < lens[:, None]
    mask = keep.float()
    mask[3, 10:60] = 0
    uids = np.array([f"prompt-{i // G}" for i in range(B)])[torch.randperm(B, generator=g).numpy()]
    rewards = torch.zeros(B, A)
    rewards[torch.arange(B), lens - 1] = torch.randn(B, generator=g)
    adv, ret = skyrl_b200.compute_rloo_outcome_advantage(rewards.to(DEV), keep.long().to(DEV), uids)
    want_adv, _ = O.rloo_advantage(rewards, keep.long(), uids)
    assert_close(adv, want_adv, what="RLOO advantages")
    assert torch.equal(adv.cpu() == 0, want_adv == 0) and adv.data_ptr() == ret.data_ptr()

    N = int(lens.sum())
    gen = torch.Generator(device=DEV).manual_seed(55)
    hidden, weight, labels = _head(N, H, V, gen)
    # a large entropy coefficient so that the entropy-gradient term is a visible share of d_hidden
    cfg = O.make_cfg(policy_loss_type="regular", loss_reduction="token_mean", use_kl_loss=False,
                     use_entropy_loss=True, entropy_loss_coef=0.5)
    cu, (advp, maskp) = skyrl_b200.pack_by_lengths(lens, adv.cpu(), mask)
    cu, advp, maskp = cu.to(DEV), advp.to(DEV), maskp.to(DEV)
    lp0 = _current_log_probs(hidden, weight, labels, (N,), cfg, T, loss_mask=maskp, cu_seqlens=cu)
    olpp = lp0 + 0.05 * torch.randn((N,), generator=gen, device=DEV)
    out = skyrl_b200.policy_update_step(hidden, weight, labels, olpp, advp, cfg, loss_mask=maskp, temperature=T,
                                        cu_seqlens=cu)
    assert torch.allclose(out.log_probs, lp0, atol=2e-5, rtol=0) and out.log_probs.shape == (N,)
    olp = torch.zeros(B, A)
    olp[keep] = olpp.cpu()
    print("config 5:", check_step(out, hidden, weight, labels, olp, adv, cfg, mask, None, T, keep=keep))


@pytest.mark.parametrize("kind,red", [("sapo", "sequence_mean"), ("cispo", "token_mean"),
                                      ("importance_sampling", "token_mean"), ("cross_entropy", "token_mean")])
def test_fused_step_runs_the_rest_of_the_registry(kind, red):
    """sapo / cispo / importance_sampling / cross_entropy through the ONE-call fused step (the run-time-typed
    loss_flat_kernel, not the per-type instantiations the registry functions launch) against the oracle composed as
    the worker composes it: RL losses add the KL and entropy terms, cross_entropy is the bare masked sum
    (worker.py:810-877); the last two reduce with a masked SUM whatever `loss_reduction` says."""
    from types import SimpleNamespace as NS

    B, A, H, V, T = 6, 200, 64, 1200, 0.9
    g = torch.Generator().manual_seed(31)
    hidden = torch.randn(B, A, H, generator=g).bfloat16()
    weight = (torch.randn(V, H, generator=g) * (1.5 / H ** 0.5)).bfloat16()
    labels = torch.randint(0, V, (B, A), generator=g)
    lens = torch.randint(A // 2, A + 1, (B,), generator=g)
    mask = (torch.arange(A)[None] < lens[:, None]).float()
    cfg = O.make_cfg(policy_loss_type=kind, loss_reduction=red, use_kl_loss=True, use_entropy_loss=True,
                     sapo=NS(tau_pos=1.0, tau_neg=1.05), cispo=NS(cispo_eps_clip_low=0.1, cispo_eps_clip_high=0.15))
    h = hidden.float().reshape(B * A, H).requires_grad_(True)
    w = weight.float().requires_grad_(True)
    lp_ref, ent_ref = O.head_logprobs_entropy(h, w, labels.reshape(-1), T)
    lp_ref, ent_ref = lp_ref.view(B, A), ent_ref.view(B, A)
    olp = lp_ref.detach() + 0.1 * torch.randn(B, A, generator=g)
    base = lp_ref.detach() + 0.1 * torch.randn(B, A, generator=g)
    adv = torch.randn(B, A, generator=g)
    final_ref, st_ref = O.loss_assembly(lp_ref, olp, adv, cfg, mask, base, ent_ref)
    final_ref.backward()
    out = skyrl_b200.policy_update_step(hidden.to(DEV), weight.to(DEV), labels.to(DEV), olp.to(DEV), adv.to(DEV), cfg,
                                        loss_mask=mask.to(DEV), base_log_probs=base.to(DEV), temperature=T)
    for k in ("final_loss", "policy_loss"):
        assert out.status[k] == pytest.approx(st_ref[k], rel=2e-3, abs=2e-4), (kind, k)
    # the oracle clips on ITS log-probs (1e-5 away from the kernel's): allow two tokens to sit on a clip boundary
    assert abs(out.status["loss_metrics/clip_ratio"] - st_ref["loss_metrics/clip_ratio"]) <= 2.5 / mask.sum().item()
    if kind != "cross_entropy":
        assert out.status["policy_kl"] == pytest.approx(st_ref["policy_kl"], rel=2e-2, abs=1e-4)
        assert out.status["policy_entropy"] == pytest.approx(st_ref["policy_entropy"], rel=2e-3)
    rel = lambda a, b: float((a.double().cpu() - b.double()).norm() / b.double().norm())
    assert rel(out.d_hidden.float(), h.grad) < 2e-2
    assert rel(out.d_weight, w.grad) < 2e-2


@pytest.mark.parametrize("kind", ["k1", "abs", "k2", "k3"])
def test_kl_loss_term_with_the_gradient_flag(kind):
    """The reference's KL-loss term carries no gradient (compute_approx_kl runs under no_grad, SURVEY trap 6) and that
    is the default here; `kl_has_grad=True` is the opt-in "fixed" behaviour: d(kl_coef * mean_b masked_mean_t kl)/d lp,
    including the clamp pass-through rules of k3, against autograd."""
    B, A, coef = 5, 300, 0.3
    g = torch.Generator().manual_seed(17)
    lp = -3.0 * torch.rand(B, A, generator=g)
    olp = lp + 0.1 * torch.randn(B, A, generator=g)
    base = lp + 0.3 * torch.randn(B, A, generator=g)
    base[0, :4] += torch.tensor([25.0, -25.0, 2.5, -2.5])  # beyond the +-20 clamp of x, and exp(x) - x - 1 > 10
    adv = torch.randn(B, A, generator=g)
    lens = torch.randint(1, A + 1, (B,), generator=g)
    lens[0] = A
    mask = (torch.arange(A)[None] < lens[:, None]).float()
    cfg = O.make_cfg(policy_loss_type="regular", loss_reduction="token_mean", use_kl_loss=True,
                     kl_estimator_type=kind, kl_loss_coef=coef)
    x = lp.clone().requires_grad_(True)
    pl, _ = O.policy_loss(x, olp, adv, cfg, mask)
    d = x - base
    if kind == "k3":
        z = (-d).clamp(-20.0, 20.0)
        kl = (z.exp() - z - 1.0).clamp(-10.0, 10.0)
    else:
        kl = {"k1": d, "abs": d.abs(), "k2": 0.5 * d * d}[kind]
    klm = O.masked_mean(kl * mask, mask, dim=-1).mean()
    (pl + coef * klm).backward()

    y = lp.to(DEV).requires_grad_(True)
    loss, st = skyrl_b200.policy_update_loss(y, olp.to(DEV), adv.to(DEV), cfg, loss_mask=mask.to(DEV),
                                             base_log_probs=base.to(DEV), kl_has_grad=True)
    loss.backward()
    assert_close(loss, pl + coef * klm, what="final loss")
    assert st["policy_kl"] == pytest.approx(klm.item(), rel=1e-5, abs=1e-5)
    assert_close(y.grad, x.grad, floor=1e-4, what=f"d_log_probs with the {kind} KL gradient")  # KL part ~ 6e-5
    # and the default: the KL term moves the value only
    y0 = lp.to(DEV).requires_grad_(True)
    loss0, _ = skyrl_b200.policy_update_loss(y0, olp.to(DEV), adv.to(DEV), cfg, loss_mask=mask.to(DEV),
                                             base_log_probs=base.to(DEV))
    loss0.backward()
    x0 = lp.clone().requires_grad_(True)
    O.policy_loss(x0, olp, adv, cfg, mask)[0].backward()
    assert_close(loss0, loss)
    assert_close(y0.grad, x0.grad, floor=1e-4, what="d_log_probs without the KL gradient")


def test_sharded_gae_equals_the_batch_gae():
    """GAE whitening statistics are global (the trainer computes advantages before sharding, SURVEY trap 12).
    dist.gae_advantage_sharded = scan -> all-reduce of the fp64 moments {sum m, sum m*x, sum m*x^2} -> whitening:
    with world size 1 it must be compute_gae_advantage_return bit for bit, and two half-batches whose moments are
    added (what the all-reduce does) must reproduce the whole-batch result."""
    from skyrl_b200 import _lib
    from skyrl_b200 import dist as sdist

    B, A = 6, 3000
    g = torch.Generator().manual_seed(23)
    lens = torch.randint(A // 2, A + 1, (B,), generator=g)
    mask = (torch.arange(A)[None] < lens[:, None]).float().to(DEV)
    r = torch.zeros(B, A)
    r[torch.arange(B), lens - 1] = torch.randn(B, generator=g)
    r, v = r.to(DEV), torch.randn(B, A, generator=g).to(DEV)
    a0, r0 = skyrl_b200.compute_gae_advantage_return(r, v, mask, 1.0, 0.95)
    a1, r1 = sdist.gae_advantage_sharded(r, v, mask, 1.0, 0.95)
    assert torch.equal(a0, a1) and torch.equal(r0, r1)

    lib = _lib.load()
    halves, moments = [], []
    for rows in (slice(0, 2), slice(2, B)):  # uneven shards
        rr, vv, mm = r[rows].contiguous(), v[rows].contiguous(), mask[rows].contiguous()
        b = rr.shape[0]
        adv, ret = torch.empty_like(rr), torch.empty_like(rr)
        mom = torch.zeros(3, dtype=torch.float64, device=DEV)
        ws = torch.empty(lib.skyb200_gae_workspace(b), dtype=torch.uint8, device=DEV)
        _lib.check(lib.skyb200_gae_scan(_lib.ptr(rr), _lib.ptr(vv), _lib.ptr(mm), _lib.MASK_F32, _lib.ptr(adv),
                                        _lib.ptr(ret), _lib.ptr(mom), _lib.ptr(ws), ws.numel(), b, A, 1.0, 0.95,
                                        _lib.stream_ptr()), "skyb200_gae_scan")
        halves.append((adv, ret))
        moments.append(mom)
    total = moments[0] + moments[1]
    stats = torch.empty(4, dtype=torch.float32, device=DEV)
    for adv, _ in halves:
        _lib.check(lib.skyb200_whiten_from_moments(_lib.ptr(adv), _lib.ptr(total), _lib.ptr(stats), adv.numel(),
                                                   _lib.stream_ptr()), "skyb200_whiten_from_moments")
    assert stats[0].item() == mask.sum().item()
    assert torch.equal(torch.cat([h[1] for h in halves]), r0)            # returns do not depend on the statistics
    assert_close(torch.cat([h[0] for h in halves]), a0, rel=1e-6, what="advantages from summed shard moments")


@pytest.mark.parametrize("A", [1000, 1001])
def test_gspo_fused_step_with_a_response_length_off_the_tile_grid(A):
    """The padded response length is whatever the longest response of the batch was.  GSPO chunks are whole
    sequences, so with A % 256 != 0 the chunks start off the 256-row tile grid (here three chunks of whole sequences,
    the later ones at token offsets that are no multiple of 256) and end in a partial tile; A = 1001 also takes the
    scalar paths."""
    from skyrl_b200 import _lib
    import ctypes

    B, H, V, T = 60, 64, 512, 1.0
    plan = (ctypes.c_int64 * 3)()
    _lib.load().skyb200_policy_step_plan(B, A, 0, H, V, _lib.LOSS_TYPES["gspo"], 1, plan)
    assert plan[1] % A == 0 and plan[1] % 256 != 0 and plan[2] >= 3, list(plan)
    g = torch.Generator().manual_seed(77)
    hidden = torch.randn(B, A, H, generator=g).bfloat16()
    weight = (torch.randn(V, H, generator=g) * (1.5 / H ** 0.5)).bfloat16()
    labels = torch.randint(0, V, (B, A), generator=g)
    lens = torch.randint(A // 2, A + 1, (B,), generator=g)
    mask = (torch.arange(A)[None] < lens[:, None]).float()
    cfg = O.make_cfg(policy_loss_type="gspo", loss_reduction="sequence_mean", use_kl_loss=True, use_entropy_loss=True,
                     eps_clip_low=0.02, eps_clip_high=0.02)
    h = hidden.float().reshape(B * A, H).requires_grad_(True)
    w = weight.float().requires_grad_(True)
    lp_ref, ent_ref = O.head_logprobs_entropy(h, w, labels.reshape(-1), T)
    lp_ref, ent_ref = lp_ref.view(B, A), ent_ref.view(B, A)
    olp = lp_ref.detach() + 0.03 * torch.randn(B, 1, generator=g) + 0.01 * torch.randn(B, A, generator=g)
    base = lp_ref.detach() + 0.1 * torch.randn(B, A, generator=g)
    adv = torch.randn(B, A, generator=g)
    final_ref, st_ref = O.loss_assembly(lp_ref, olp, adv, cfg, mask, base, ent_ref)
    final_ref.backward()
    out = skyrl_b200.policy_update_step(hidden.to(DEV), weight.to(DEV), labels.to(DEV), olp.to(DEV), adv.to(DEV), cfg,
                                        loss_mask=mask.to(DEV), base_log_probs=base.to(DEV), temperature=T)
    assert (out.log_probs.cpu() - lp_ref.detach()).abs().max().item() < 1e-3
    assert (out.entropy.cpu() - ent_ref.detach()).abs().max().item() < 1e-3
    assert 0.0 < st_ref["loss_metrics/clip_ratio"] < 1.0
    for k in ("final_loss", "policy_loss", "policy_entropy"):
        assert out.status[k] == pytest.approx(st_ref[k], rel=2e-3, abs=2e-4), k
    # a whole sequence flips its clip state if its ratio sits within 1e-5 of a boundary: allow one
    assert abs(out.status["loss_metrics/clip_ratio"] - st_ref["loss_metrics/clip_ratio"]) <= 1.2 * A / mask.sum().item()
    rel = lambda a, b: float((a.double().cpu() - b.double()).norm() / b.double().norm())
    assert rel(out.d_hidden.float(), h.grad) < 2e-2
    assert rel(out.d_weight, w.grad) < 2e-2


def test_fused_head_policy_on_a_tiny_qwen2():
    """FusedHeadPolicy (the drop-in for HFModelWrapper.forward) end to end on the GPU: bf16 trunk of a tiny random
    Qwen2 -> CUDA fused head.  Log-probs / entropy against the oracle head on the trunk's own hidden states (the
    parity definition: fp32 logits from the bf16 operands, 1e-3 abs); gradients reach the lm_head and the trunk."""
    from transformers import Qwen2Config, Qwen2ForCausalLM

    torch.manual_seed(0)
    cfg = Qwen2Config(vocab_size=1000, hidden_size=64, intermediate_size=128, num_hidden_layers=2,
                      num_attention_heads=4, num_key_value_heads=2, max_position_embeddings=128,
                      tie_word_embeddings=False)
    model = Qwen2ForCausalLM(cfg).to(device=DEV, dtype=torch.bfloat16)
    policy = skyrl_b200.FusedHeadPolicy(model)
    B, S, A, T = 4, 40, 17, 0.8
    sequences = torch.randint(0, 1000, (B, S), device=DEV)
    am = torch.ones(B, S, dtype=torch.long, device=DEV)
    am[0, :5] = 0
    am[3, -3:] = 0
    lp, out = policy(sequences, A, am, temperature=T, return_output=True, compute_entropy=True)
    assert lp.shape == (B, A) and out["entropy"].shape == (B, S) and lp.dtype == torch.float32
    with torch.no_grad():
        pos = am.cumsum(-1) - 1
        pos.masked_fill_(am == 0, 1)
        hidden = model.model(sequences, attention_mask=am, position_ids=pos).last_hidden_state
    sl = slice(S - A - 1, S - 1)
    lp_ref, ent_ref = O.head_logprobs_entropy(hidden[:, sl].reshape(B * A, -1).float().cpu(),
                                              model.lm_head.weight.detach().float().cpu(),
                                              torch.roll(sequences, -1, 1)[:, sl].reshape(-1).cpu(), T)
    assert (lp.detach().cpu() - lp_ref.view(B, A)).abs().max().item() < 1e-3
    want_ent = ent_ref.view(B, A) * am[:, sl].cpu().float()
    assert (out["entropy"][:, sl].detach().cpu() - want_ent).abs().max().item() < 1e-3
    assert torch.count_nonzero(out["entropy"][:, :S - A - 1]) == 0
    (lp.mean() - 0.01 * out["entropy"][:, sl].mean()).backward()
    gw, gt = model.lm_head.weight.grad, model.model.layers[0].mlp.up_proj.weight.grad
    assert gw is not None and gt is not None and torch.isfinite(gw).all() and torch.isfinite(gt).all()
    assert gw.abs().max().item() > 0 and gt.abs().max().item() > 0


@pytest.mark.parametrize("fused", [False, True])
def test_sharply_peaked_logits(fused):
    """A trained policy is nearly deterministic on most tokens: logits with std ~13 (std 8 / T = 0.6), a third of the
    labels at the arg-max (log-prob ~ 0, entropy ~ 1e-3) and the rest far in the tail (log-prob ~ -40).  The online
    softmax, the entropy as a difference of two large numbers and both backward paths (fp32 recompute, and the fused
    step's bf16 logits store) against the fp32 oracle."""
    N, H, V, T = 768, 256, 5000, 0.6
    g = torch.Generator().manual_seed(9)
    hidden = torch.randn(N, H, generator=g).bfloat16()
    weight = (torch.randn(V, H, generator=g) * (8.0 / H ** 0.5)).bfloat16()
    labels = torch.randint(0, V, (N,), generator=g)
    labels[: N // 3] = (hidden[: N // 3].float() @ weight.float().t()).argmax(-1)
    h = hidden.float().requires_grad_(True)
    w = weight.float().requires_grad_(True)
    lp_ref, ent_ref = O.head_logprobs_entropy(h, w, labels, T)
    assert lp_ref[: N // 3].max().item() > -0.05 and lp_ref.min().item() < -25.0 and ent_ref.median().item() < 1.0
    rel = lambda a, b: float((a.double().cpu() - b.double()).norm() / b.double().norm())
    if not fused:
        g_lp, g_ent = torch.randn(N, generator=g), 0.3 * torch.randn(N, generator=g)
        ((lp_ref * g_lp).sum() + (ent_ref * g_ent).sum()).backward()
        hd, wd = hidden.to(DEV).requires_grad_(True), weight.to(DEV).requires_grad_(True)
        lp, ent = skyrl_b200.fused_lm_head(hd, wd, labels.to(DEV), temperature=T)
        ((lp * g_lp.to(DEV)).sum() + (ent * g_ent.to(DEV)).sum()).backward()
        d_hidden, d_weight = hd.grad.float(), wd.grad.float()
    else:
        B, A = 3, N // 3
        cfg = O.make_cfg(policy_loss_type="regular", loss_reduction="token_mean", use_kl_loss=False,
                         use_entropy_loss=True, entropy_loss_coef=0.3)
        olp = lp_ref.detach().view(B, A) + 0.05 * torch.randn(B, A, generator=g)
        adv = torch.randn(B, A, generator=g)
        final, _ = O.loss_assembly(lp_ref.view(B, A), olp, adv, cfg, torch.ones(B, A), None, ent_ref.view(B, A))
        final.backward()
        out = skyrl_b200.policy_update_step(hidden.to(DEV), weight.to(DEV), labels.to(DEV).view(B, A), olp.to(DEV),
                                            adv.to(DEV), cfg, loss_mask=torch.ones(B, A, device=DEV), temperature=T)
        lp, ent, d_hidden, d_weight = out.log_probs.reshape(-1), out.entropy.reshape(-1), out.d_hidden.float(), out.d_weight
        assert out.status["final_loss"] == pytest.approx(final.item(), rel=2e-3, abs=2e-4)
    assert (lp.detach().cpu() - lp_ref.detach()).abs().max().item() < 1e-3
    assert (ent.detach().cpu() - ent_ref.detach()).abs().max().item() < 1e-3
    # Both backward paths hold the same bound at ANY logit scale: the autograd path recomputes the logits tile in
    # fp32; the fused step reads back slab-relative log2-logits (EpiLseStoreRel), whose bf16 rounding error vanishes
    # on the entries that carry the probability mass (|z| reaches ~70 here; rounding the raw logit instead would
    # cost up to 28 % per probability and ~6 % on d_hidden - the round-1 kernel, bound 0.12).  The reference's own
    # bf16-logits arithmetic sits at 2.7 % on this input (CPU emulation).
    assert rel(d_hidden, h.grad) < 1e-2 and rel(d_weight, w.grad) < 1e-2
    # every d(logits) row sums to ~0 (softmax rows sum to one): d_weight summed over the vocabulary cancels
    dw = d_weight.double().cpu()
    assert float((dw.sum(0).abs() / dw.abs().sum(0)).max()) < 2e-3


@pytest.mark.parametrize("std,T", [(1.5, 1.0), (4.0, 1.0), (8.0, 0.6), (16.0, 1.0)])
def test_fused_backward_error_does_not_grow_with_the_logit_scale(std, T):
    """The fused step's backward (bf16 tile store -> transform -> two GEMMs) against fp32 autograd of the oracle for
    logit scales from random-init (std 1.5) to far beyond any trained head (std 16, |z| > 100): relative L2 error of
    d_hidden / d_weight stays at the bf16-operand floor, forward outputs within 1e-3, rows of d(logits) sum to zero.
    V is not a slab multiple and the last token tile is ragged."""
    B, A, H, V = 3, 400, 256, 5000
    g = torch.Generator().manual_seed(11)
    hidden = torch.randn(B * A, H, generator=g).bfloat16()
    weight = (torch.randn(V, H, generator=g) * (std / H ** 0.5)).bfloat16()
    labels = torch.randint(0, V, (B * A,), generator=g)
    labels[: B * A // 3] = (hidden[: B * A // 3].float() @ weight.float().t()).argmax(-1)
    h = hidden.float().requires_grad_(True)
    w = weight.float().requires_grad_(True)
    lp_ref, ent_ref = O.head_logprobs_entropy(h, w, labels, T)
    cfg = O.make_cfg(policy_loss_type="regular", loss_reduction="token_mean", use_kl_loss=False,
                     use_entropy_loss=True, entropy_loss_coef=0.01)
    olp = lp_ref.detach().view(B, A) + 0.05 * torch.randn(B, A, generator=g)
    adv = torch.randn(B, A, generator=g)
    mask = torch.ones(B, A)
    mask[1, 300:] = 0
    final, _ = O.loss_assembly(lp_ref.view(B, A), olp, adv, cfg, mask, None, ent_ref.view(B, A))
    final.backward()
    out = skyrl_b200.policy_update_step(hidden.to(DEV), weight.to(DEV), labels.to(DEV).view(B, A), olp.to(DEV),
                                        adv.to(DEV), cfg, loss_mask=mask.to(DEV), temperature=T)
    rel = lambda a, b: float((a.double().cpu() - b.double()).norm() / b.double().norm())
    assert (out.log_probs.cpu().view(-1) - lp_ref.detach()).abs().max().item() < 1e-3
    assert (out.entropy.cpu().view(-1) - ent_ref.detach()).abs().max().item() < 1e-3
    assert out.status["final_loss"] == pytest.approx(final.item(), rel=2e-3, abs=2e-4)
    r_dh, r_dw = rel(out.d_hidden.float(), h.grad), rel(out.d_weight, w.grad)
    dw = out.d_weight.double().cpu()
    colsum = float((dw.sum(0).abs() / dw.abs().sum(0)).max())
    print(f"std {std} T {T}: rel d_hidden {r_dh:.2e} d_weight {r_dw:.2e} colsum {colsum:.2e}")
    assert r_dh < 6e-3 and r_dw < 6e-3 and colsum < 2e-3


@pytest.mark.parametrize("kw", [
    dict(tis_ratio_type="token"),
    dict(tis_ratio_type="sequence", sequence_mask_metric="geometric", geo_mask_low=0.98, geo_mask_high=1.02),
    dict(sequence_mask_metric="product", outlier_token_is_threshold_low=0.3, outlier_token_is_threshold_high=3.0),
])
@pytest.mark.parametrize("fused", [True, False])
def test_step_applies_the_off_policy_correction(kw, fused):
    """policy_update_step with rollout_logprobs and config.off_policy_correction: TIS weights, outlier-token and
    sequence masks change the loss, its mask and the gradients exactly as in the registry losses
    (ppo_utils.py:577-583), clip_ratio keeps the uncorrected mask, and the correction's metrics are reported."""
    B, A, H, V, T = 6, 300, 64, 1500, 1.0
    g = torch.Generator().manual_seed(21)
    hidden = torch.randn(B * A, H, generator=g).bfloat16()
    weight = (torch.randn(V, H, generator=g) * (1.5 / H ** 0.5)).bfloat16()
    labels = torch.randint(0, V, (B, A), generator=g)
    lens = torch.randint(A // 3, A + 1, (B,), generator=g)
    mask = (torch.arange(A)[None] < lens[:, None]).float()
    h = hidden.float().requires_grad_(True)
    w = weight.float().requires_grad_(True)
    lp_ref, ent_ref = O.head_logprobs_entropy(h, w, labels.reshape(-1), T)
    olp = lp_ref.detach().view(B, A) + 0.2 * torch.randn(B, A, generator=g)
    rlp = olp + 0.3 * torch.randn(B, A, generator=g) * (torch.rand(B, 1, generator=g) < 0.7)
    rlp[0] = olp[0] + 0.001 * torch.randn(A, generator=g)   # one sequence inside every sequence-mask window
    adv = torch.randn(B, A, generator=g)
    opc = NS(tis_ratio_type=None, sequence_mask_metric=None, token_tis_ratio_clip_high=2.0,
             sequence_tis_ratio_clip_high=5.0, geo_mask_low=0.99, geo_mask_high=1.01, product_mask_low=0.5,
             product_mask_high=2.0, outlier_token_is_threshold_low=None, outlier_token_is_threshold_high=None)
    for k, v in kw.items():
        setattr(opc, k, v)
    cfg = O.make_cfg(policy_loss_type="regular", loss_reduction="token_mean", use_kl_loss=False,
                     off_policy_correction=opc)
    loss_ref, met_ref = O.policy_loss_with_correction(lp_ref.view(B, A), olp, adv, cfg, mask, rlp)
    loss_ref.backward()
    out = skyrl_b200.policy_update_step(hidden.to(DEV), weight.to(DEV), labels.to(DEV), olp.to(DEV), adv.to(DEV), cfg,
                                        loss_mask=mask.to(DEV), temperature=T, rollout_logprobs=rlp.to(DEV),
                                        fused=fused)
    assert out.status["policy_loss"] == pytest.approx(loss_ref.item(), rel=1e-4, abs=1e-6)
    assert out.status["loss_metrics/clip_ratio"] == pytest.approx(met_ref["clip_ratio"], rel=1e-5, abs=1e-7)
    for k, v in met_ref.items():
        if k != "clip_ratio":
            assert out.status["loss_metrics/" + k] == pytest.approx(v, rel=1e-4, abs=1e-5), k
    rel = lambda a, b: float((a.double().cpu() - b.double()).norm() / b.double().norm())
    assert rel(out.d_hidden.float(), h.grad) < 1e-2 and rel(out.d_weight, w.grad) < 1e-2
    # and without rollout_logprobs the correction is off: a different loss
    plain = skyrl_b200.policy_update_step(hidden.to(DEV), weight.to(DEV), labels.to(DEV), olp.to(DEV), adv.to(DEV),
                                          cfg, loss_mask=mask.to(DEV), temperature=T, fused=fused)
    assert abs(plain.status["policy_loss"] - out.status["policy_loss"]) > 1e-6


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (NCCL data-parallel group)")
def test_data_parallel_step_over_nccl_two_ranks():
    """N-GPU == 1-GPU on the same batch: whole prompt groups per rank -> local GRPO advantages, per-rank fused step,
    packed-stats and fp32 d_weight all-reduce (mean of per-rank means, the reference's semantics) against the
    single-GPU step on the whole batch; GAE with the three whitening moments all-reduced against the whole-batch GAE."""
    import os
    import subprocess
    import sys

    script = os.path.join(os.path.dirname(__file__), "dp_worker.py")
    env = dict(os.environ, PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", script], env=env, capture_output=True,
                       text=True, timeout=280)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "DP_OK" in r.stdout


@pytest.mark.parametrize("remote", ["red", "tma"])
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (NCCL data-parallel group)")
def test_fused_dweight_reduce_scatter_over_nccl_two_ranks(remote):
    """The d_weight GEMM whose epilogue reduce-scatters the gradient over NVLink peer memory (TMA reduce-add for the
    rank's own rows, fp32 vector reductions on the peers' shards) == NCCL's mean reduce-scatter of the local
    gradients: single chunk, several chunks with masked rows, accumulated micro-batches, repeated optimizer steps."""
    import os
    import subprocess
    import sys

    script = os.path.join(os.path.dirname(__file__), "dws_worker.py")
    env = dict(os.environ, PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
               SKYB200_TEST_DWS_REMOTE=remote)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29543", script], env=env, capture_output=True,
                       text=True, timeout=280)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "DWS_OK" in r.stdout


@pytest.mark.parametrize("dead", [0, 1, 2])
def test_gspo_chunk_whose_every_row_is_dead(dead):
    """GSPO clips whole sequences: a sequence whose ratio left the clip range on the side its advantage clips carries
    no gradient at all, and with long sequences that is a whole chunk of all-zero d(logits) rows - the compacted
    backward GEMMs then run over ZERO live tiles / K blocks (first chunk: plain GEMM over zeros, it must still
    initialise d_weight).  Three sequences of 20000 tokens = three chunks; each of them dead in turn."""
    from skyrl_b200 import _lib
    import ctypes

    B, A, H, V, T = 3, 20000, 64, 512, 1.0
    plan = (ctypes.c_int64 * 3)()
    _lib.load().skyb200_policy_step_plan(B, A, 0, H, V, _lib.LOSS_TYPES["gspo"], 1, plan)
    assert plan[1] == A and plan[2] == 3, list(plan)                       # one sequence per chunk
    g = torch.Generator().manual_seed(13)
    hidden = torch.randn(B, A, H, generator=g).bfloat16()
    weight = (torch.randn(V, H, generator=g) * (1.5 / H ** 0.5)).bfloat16()
    labels = torch.randint(0, V, (B, A), generator=g)
    lens = torch.tensor([A, A - 777, A - 5000])
    mask = (torch.arange(A)[None] < lens[:, None]).float()
    cfg = O.make_cfg(policy_loss_type="gspo", loss_reduction="sequence_mean", use_kl_loss=False,
                     eps_clip_low=0.05, eps_clip_high=0.05)
    h = hidden.float().reshape(B * A, H).requires_grad_(True)
    w = weight.float().requires_grad_(True)
    lp_ref, ent_ref = O.head_logprobs_entropy(h, w, labels.reshape(-1), T)
    lp_ref, ent_ref = lp_ref.view(B, A), ent_ref.view(B, A)
    off = torch.full((B, 1), 0.01)
    off[dead] = 0.2                                                        # ratio e^0.2 = 1.22 > 1.05 with A > 0: clipped
    olp = lp_ref.detach() - off + 0.002 * torch.randn(B, A, generator=g)
    adv = torch.tensor([[0.8], [1.1], [0.6]]).expand(B, A).contiguous()
    final_ref, st_ref = O.loss_assembly(lp_ref, olp, adv, cfg, mask, None, ent_ref)
    final_ref.backward()
    assert h.grad.view(B, A, H)[dead].abs().max().item() == 0.0            # the oracle agrees: no gradient there
    out = skyrl_b200.policy_update_step(hidden.to(DEV), weight.to(DEV), labels.to(DEV), olp.to(DEV), adv.to(DEV), cfg,
                                        loss_mask=mask.to(DEV), temperature=T)
    assert out.status["final_loss"] == pytest.approx(st_ref["final_loss"], rel=2e-3, abs=2e-4)
    assert torch.count_nonzero(out.d_hidden.view(B, A, H)[dead]) == 0
    rel = lambda a, b: float((a.double().cpu() - b.double()).norm() / b.double().norm())
    assert rel(out.d_hidden.float(), h.grad) < 2e-2
    assert rel(out.d_weight, w.grad) < 2e-2
