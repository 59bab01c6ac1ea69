"""The whole step (head fwd -> loss assembly -> head bwd) against the oracle composed the way the reference
worker composes it (worker.py:787-889), at a size the CPU oracle finishes in seconds."""
import numpy as np
import pytest
import torch

import skyrl_b200
from oracle import policy_oracle as O
from conftest import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _batch(B, A, H, V, seed):
    g = torch.Generator().manual_seed(seed)
    hidden = torch.randn(B, A, H, generator=g).bfloat16()
    weight = (torch.randn(V, H, generator=g) * (1.5 / H ** 0.5)).bfloat16()
    labels = torch.randint(0, V, (B, A), generator=g)
    lens = torch.randint(A // 2, A + 1, (B,), generator=g)
    mask = (torch.arange(A)[None] < lens[:, None]).float()
    return g, hidden, weight, labels, mask


@pytest.mark.parametrize("kw", [
    dict(policy_loss_type="regular", loss_reduction="token_mean", use_kl_loss=True, use_entropy_loss=False),
    dict(policy_loss_type="dual_clip", loss_reduction="token_mean", use_kl_loss=False, use_entropy_loss=True,
         eps_clip_high=0.28, clip_ratio_c=10.0),
    dict(policy_loss_type="gspo", loss_reduction="sequence_mean", use_kl_loss=True, use_entropy_loss=True,
         eps_clip_low=0.05, eps_clip_high=0.05),
])
@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("shape", [(8, 96, 128, 2000), (5, 1000, 64, 1500)])
def test_step_matches_oracle(kw, fused, shape):
    B, A, H, V = shape
    T = 0.9
    g, hidden, weight, labels, mask = _batch(B, A, H, V, 42)
    cfg = O.make_cfg(**kw)
    h = hidden.float().reshape(B * A, H).requires_grad_(True)
    w = weight.float().requires_grad_(True)
    lp_ref, ent_ref = O.head_logprobs_entropy(h, w, labels.reshape(-1), T)
    lp_ref, ent_ref = lp_ref.view(B, A), ent_ref.view(B, A)
    olp = lp_ref.detach() + 0.05 * torch.randn(B, A, generator=g)
    base = lp_ref.detach() + 0.1 * torch.randn(B, A, generator=g)
    adv = torch.randn(B, A, generator=g)
    final_ref, st_ref = O.loss_assembly(lp_ref, olp, adv, cfg, mask, base, ent_ref)
    final_ref.backward()

    out = skyrl_b200.policy_update_step(hidden.to(DEV), weight.to(DEV), labels.to(DEV), olp.to(DEV), adv.to(DEV),
                                        cfg, loss_mask=mask.to(DEV), base_log_probs=base.to(DEV), temperature=T,
                                        fused=fused)
    assert (out.log_probs.cpu() - lp_ref.detach()).abs().max().item() < 1e-3
    for k in ("final_loss", "policy_loss", "policy_entropy", "loss_metrics/clip_ratio"):
        assert out.status[k] == pytest.approx(st_ref[k], rel=2e-3, abs=2e-4), k
    if cfg.use_kl_loss:
        assert out.status["policy_kl"] == pytest.approx(st_ref["policy_kl"], rel=2e-2, abs=1e-4)
    rel = lambda a, b: float((a.double().cpu() - b.double()).norm() / b.double().norm())
    assert rel(out.d_hidden.float(), h.grad) < 2e-2
    assert rel(out.d_weight, w.grad) < 2e-2

    # the same step through autograd + the registry loss (the reference-shaped call path) gives the same numbers
    hd = hidden.to(DEV).reshape(B * A, H).requires_grad_(True)
    wd = weight.to(DEV).requires_grad_(True)
    lp, ent = skyrl_b200.fused_lm_head(hd, wd, labels.to(DEV).reshape(-1), temperature=T)
    loss, st = skyrl_b200.policy_update_loss(lp.view(B, A), olp.to(DEV), adv.to(DEV), cfg, loss_mask=mask.to(DEV),
                                             base_log_probs=base.to(DEV), entropy=ent.view(B, A))
    loss.backward()
    assert st["final_loss"] == pytest.approx(out.status["final_loss"], rel=1e-5, abs=1e-6)
    # fused keeps bf16 logits for the backward, the autograd path recomputes them in fp32: both within tolerance
    assert rel(hd.grad.float(), out.d_hidden.float().cpu()) < (1e-2 if fused else 1e-3)


def test_fused_step_multi_chunk_entropy_mask_and_accumulate():
    """Several chunks (24576 tokens = 18944 + a ragged one), an entropy mask, entropy gradient and d_weight accumulation."""
    B, A, H, V, T = 6, 4096, 64, 1024, 1.0
    g, hidden, weight, labels, mask = _batch(B, A, H, V, 5)
    emask = (torch.rand(B, A, generator=g) > 0.1).float()
    cfg = O.make_cfg(policy_loss_type="regular", loss_reduction="sequence_mean", use_kl_loss=True,
                     use_entropy_loss=True, entropy_loss_coef=0.05)
    h = hidden.float().reshape(B * A, H).requires_grad_(True)
    w = weight.float().requires_grad_(True)
    lp_ref, ent_ref = O.head_logprobs_entropy(h, w, labels.reshape(-1), T)
    lp_ref, ent_ref = lp_ref.view(B, A), ent_ref.view(B, A) * emask
    olp = lp_ref.detach() + 0.05 * torch.randn(B, A, generator=g)
    base = lp_ref.detach() + 0.1 * torch.randn(B, A, generator=g)
    adv = torch.randn(B, A, generator=g)
    final_ref, st_ref = O.loss_assembly(lp_ref, olp, adv, cfg, mask, base, ent_ref)
    final_ref.backward()
    dw = torch.ones(V, H, device=DEV)
    args = (hidden.to(DEV), weight.to(DEV), labels.to(DEV), olp.to(DEV), adv.to(DEV), cfg)
    kw = dict(loss_mask=mask.to(DEV), base_log_probs=base.to(DEV), temperature=T, entropy_mask=emask.to(DEV))
    out = skyrl_b200.policy_update_step(*args, d_weight=dw, accumulate_d_weight=True, **kw)
    rel = lambda a, b: float((a.double().cpu() - b.double()).norm() / b.double().norm())
    for k in ("final_loss", "policy_loss", "policy_entropy", "policy_kl", "loss_metrics/clip_ratio"):
        assert out.status[k] == pytest.approx(st_ref[k], rel=2e-3, abs=2e-4), k
    assert torch.equal(out.entropy.cpu() == 0, ent_ref.detach() == 0)
    assert rel(out.d_hidden.float(), h.grad) < 2e-2
    assert rel(out.d_weight - 1.0, w.grad) < 2e-2  # accumulated on top of the ones
    # loss-only evaluation (no gradients requested) gives the same status
    out2 = skyrl_b200.policy_update_step(*args, need_d_hidden=False, need_d_weight=False, **kw)
    assert out2.d_hidden is None and out2.d_weight is None
    assert out2.status["final_loss"] == pytest.approx(out.status["final_loss"], rel=1e-6)


def test_grpo_config1_slice_end_to_end():
    """BASELINE config 1 pieces chained: GRPO advantages (group 8) -> head -> regular loss, on a slice."""
    B, A, H, V = 16, 64, 896, 151936
    g, hidden, weight, labels, mask = _batch(B, A, H, V, 7)
    uids = [f"prompt-{i // 8}" for i in range(B)]
    rewards = torch.zeros(B, A)
    rewards[torch.arange(B), mask.sum(-1).long() - 1] = (torch.rand(B, generator=g) < 0.4).float()
    adv_ref, _ = O.grpo_advantage(rewards, mask, uids)
    adv, _ = skyrl_b200.compute_grpo_outcome_advantage(rewards.to(DEV), mask.to(DEV), uids)
    assert_close(adv, adv_ref)
    cfg = O.make_cfg(use_kl_loss=False)
    lp_ref, _ = O.head_logprobs_entropy(hidden.reshape(-1, H), weight, labels.reshape(-1), 1.0)
    olp = lp_ref.view(B, A) + 0.02 * torch.randn(B, A, generator=g)
    ref_loss, ref_clip = O.policy_loss(lp_ref.view(B, A), olp, adv_ref, cfg, mask)
    out = skyrl_b200.policy_update_step(hidden.to(DEV), weight.to(DEV), labels.to(DEV), olp.to(DEV), adv, cfg,
                                        loss_mask=mask.to(DEV))
    assert out.status["policy_loss"] == pytest.approx(ref_loss.item(), rel=5e-3, abs=5e-4)
    assert out.d_hidden.shape == (B * A, H) and out.d_weight.shape == (V, H)
    assert torch.isfinite(out.d_weight).all() and torch.isfinite(out.d_hidden.float()).all()


@pytest.mark.parametrize("kw", [
    dict(policy_loss_type="regular", loss_reduction="token_mean", use_kl_loss=True, use_entropy_loss=True),
    dict(policy_loss_type="dual_clip", loss_reduction="sequence_mean", use_kl_loss=True, use_entropy_loss=False),
    dict(policy_loss_type="regular", loss_reduction="seq_mean_token_sum_norm", use_kl_loss=False, max_seq_len=700),
])
def test_packed_ragged_layout_equals_padded_reference(kw):
    """Config-5 style ragged batch: the packed layout (no padding rows at all) must give the numbers of the padded
    reference computation with loss_mask = 0 on the padding, plus an interior hole in the loss mask."""
    B, A, H, V, T = 7, 600, 64, 1500, 1.0
    g = torch.Generator().manual_seed(11)
    hidden = torch.randn(B, A, H, generator=g).bfloat16()
    weight = (torch.randn(V, H, generator=g) * (1.5 / H ** 0.5)).bfloat16()
    labels = torch.randint(0, V, (B, A), generator=g)
    lens = torch.tensor([600, 37, 1, 420, 128, 599, 300])
    pad_mask = (torch.arange(A)[None] < lens[:, None]).float()
    loss_mask = pad_mask.clone()
    loss_mask[3, 100:180] = 0  # tool tokens inside a response
    cfg = O.make_cfg(**kw)
    h = hidden.float().reshape(B * A, H).requires_grad_(True)
    w = weight.float().requires_grad_(True)
    lp_ref, ent_ref = O.head_logprobs_entropy(h, w, labels.reshape(-1), T)
    lp_ref, ent_ref = lp_ref.view(B, A), ent_ref.view(B, A)
    olp = lp_ref.detach() + 0.05 * torch.randn(B, A, generator=g)
    base = lp_ref.detach() + 0.1 * torch.randn(B, A, generator=g)
    adv = torch.randn(B, A, generator=g)
    final_ref, st_ref = O.loss_assembly(lp_ref, olp, adv, cfg, loss_mask, base, ent_ref)
    final_ref.backward()

    cu, (hp, labp, olpp, advp, maskp, basep) = skyrl_b200.pack_by_lengths(lens, hidden, labels, olp, adv, loss_mask, base)
    assert cu.tolist()[-1] == int(lens.sum()) and hp.shape == (int(lens.sum()), H)
    out = skyrl_b200.policy_update_step(hp.to(DEV), weight.to(DEV), labp.to(DEV), olpp.to(DEV), advp.to(DEV), cfg,
                                        loss_mask=maskp.to(DEV), base_log_probs=basep.to(DEV), temperature=T,
                                        cu_seqlens=cu.to(DEV))
    keep = pad_mask.bool()
    assert (out.log_probs.cpu() - lp_ref.detach()[keep]).abs().max().item() < 1e-3
    for k in ("final_loss", "policy_loss", "policy_entropy", "loss_metrics/clip_ratio"):
        assert out.status[k] == pytest.approx(st_ref[k], rel=2e-3, abs=2e-4), k
    if cfg.use_kl_loss:
        assert out.status["policy_kl"] == pytest.approx(st_ref["policy_kl"], rel=2e-2, abs=1e-4)
    rel = lambda a, b: float((a.double().cpu() - b.double()).norm() / b.double().norm())
    assert rel(out.d_hidden.float(), h.grad.view(B, A, H)[keep]) < 2e-2
    assert rel(out.d_weight, w.grad) < 2e-2
    with pytest.raises(RuntimeError, match="gspo"):
        skyrl_b200.policy_update_step(hp.to(DEV), weight.to(DEV), labp.to(DEV), olpp.to(DEV), advp.to(DEV),
                                      O.make_cfg(policy_loss_type="gspo", loss_reduction="sequence_mean"),
                                      loss_mask=maskp.to(DEV), cu_seqlens=cu.to(DEV))


def test_gspo_fused_step_row_aligned_multi_chunk():
    """GSPO needs whole sequences per chunk: 7 x 4096 tokens -> two chunks of whole sequences (skyb200_policy_step_plan)."""
    import ctypes

    from skyrl_b200 import _lib

    B, A, H, V, T = 7, 4096, 64, 512, 1.0
    plan = (ctypes.c_int64 * 3)()
    _lib.load().skyb200_policy_step_plan(B, A, 0, H, V, _lib.LOSS_TYPES["gspo"], 1, plan)
    assert plan[1] % A == 0 and plan[2] == 2, list(plan)
    g, hidden, weight, labels, mask = _batch(B, A, H, V, 21)
    cfg = O.make_cfg(policy_loss_type="gspo", loss_reduction="sequence_mean", use_kl_loss=False,
                     eps_clip_low=0.02, eps_clip_high=0.02)
    h = hidden.float().reshape(B * A, H).requires_grad_(True)
    w = weight.float().requires_grad_(True)
    lp_ref, ent_ref = O.head_logprobs_entropy(h, w, labels.reshape(-1), T)
    lp_ref = lp_ref.view(B, A)
    olp = lp_ref.detach() + 0.03 * torch.randn(B, 1, generator=g) + 0.01 * torch.randn(B, A, generator=g)
    adv = torch.randn(B, A, generator=g)
    final_ref, st_ref = O.loss_assembly(lp_ref, olp, adv, cfg, mask, None, ent_ref.view(B, A))
    final_ref.backward()
    out = skyrl_b200.policy_update_step(hidden.to(DEV), weight.to(DEV), labels.to(DEV), olp.to(DEV), adv.to(DEV), cfg,
                                        loss_mask=mask.to(DEV), temperature=T)
    assert 0.0 < st_ref["loss_metrics/clip_ratio"] < 1.0  # the case actually exercises the sequence-level clip
    for k in ("final_loss", "policy_loss", "loss_metrics/clip_ratio"):
        assert out.status[k] == pytest.approx(st_ref[k], rel=2e-3, abs=2e-4), k
    rel = lambda a, b: float((a.double().cpu() - b.double()).norm() / b.double().norm())
    assert rel(out.d_hidden.float(), h.grad) < 2e-2
    assert rel(out.d_weight, w.grad) < 2e-2


def test_fused_policy_loss_autograd_node():
    """loss.backward() through the single fused node == the oracle's gradients (incl. a non-unit upstream gradient)."""
    B, A, H, V, T = 6, 128, 64, 1200, 1.0
    g, hidden, weight, labels, mask = _batch(B, A, H, V, 31)
    cfg = O.make_cfg(policy_loss_type="dual_clip", loss_reduction="token_mean", use_kl_loss=True, clip_ratio_c=3.0)
    h = hidden.float().reshape(B * A, H).requires_grad_(True)
    w = weight.float().requires_grad_(True)
    lp_ref, ent_ref = O.head_logprobs_entropy(h, w, labels.reshape(-1), T)
    olp = lp_ref.detach().view(B, A) + 0.05 * torch.randn(B, A, generator=g)
    base = lp_ref.detach().view(B, A) + 0.1 * torch.randn(B, A, generator=g)
    adv = torch.randn(B, A, generator=g)
    final_ref, st_ref = O.loss_assembly(lp_ref.view(B, A), olp, adv, cfg, mask, base, ent_ref.view(B, A))
    (3.0 * final_ref).backward()
    hd = hidden.to(DEV).requires_grad_(True)
    wd = weight.to(DEV).requires_grad_(True)
    loss, stats, lp, ent = skyrl_b200.fused_policy_loss(hd, wd, labels.to(DEV), olp.to(DEV), adv.to(DEV), cfg,
                                                        loss_mask=mask.to(DEV), base_log_probs=base.to(DEV),
                                                        temperature=T)
    (3.0 * loss).backward()
    assert loss.item() == pytest.approx(st_ref["final_loss"], rel=2e-3, abs=2e-4)
    rel = lambda a, b: float((a.double().cpu() - b.double()).norm() / b.double().norm())
    assert hd.grad.shape == hd.shape and hd.grad.dtype == torch.bfloat16
    assert rel(hd.grad.float().reshape(B * A, H), h.grad) < 2e-2
    assert rel(wd.grad.float(), w.grad) < 2e-2


@pytest.mark.parametrize("shape", [(1, 1, 8, 8), (2, 3, 16, 40), (1, 300, 24, 257), (3, 129, 40, 1000)])
@pytest.mark.parametrize("fused", [True, False])
def test_degenerate_shapes(shape, fused):
    """Tiny / odd shapes: fewer rows than one tile, vocab smaller than one slab, H = one partial K block."""
    B, A, H, V = shape
    T = 1.0
    g = torch.Generator().manual_seed(B * 1000 + A + H + V)
    hidden = torch.randn(B, A, H, generator=g).bfloat16()
    weight = (torch.randn(V, H, generator=g) * (1.0 / H ** 0.5)).bfloat16()
    labels = torch.randint(0, V, (B, A), generator=g)
    mask = torch.ones(B, A)
    cfg = O.make_cfg(policy_loss_type="regular", loss_reduction="token_mean", use_kl_loss=True)
    h = hidden.float().reshape(B * A, H).requires_grad_(True)
    w = weight.float().requires_grad_(True)
    lp_ref, ent_ref = O.head_logprobs_entropy(h, w, labels.reshape(-1), T)
    olp = lp_ref.detach().view(B, A) + 0.05 * torch.randn(B, A, generator=g)
    base = lp_ref.detach().view(B, A) + 0.1 * torch.randn(B, A, generator=g)
    adv = torch.randn(B, A, generator=g)
    final_ref, st_ref = O.loss_assembly(lp_ref.view(B, A), olp, adv, cfg, mask, base, ent_ref.view(B, A))
    final_ref.backward()
    out = skyrl_b200.policy_update_step(hidden.to(DEV), weight.to(DEV), labels.to(DEV), olp.to(DEV), adv.to(DEV), cfg,
                                        loss_mask=mask.to(DEV), base_log_probs=base.to(DEV), temperature=T,
                                        fused=fused)
    assert (out.log_probs.cpu() - lp_ref.detach().view(B, A)).abs().max().item() < 1e-3
    assert out.status["final_loss"] == pytest.approx(st_ref["final_loss"], rel=2e-3, abs=2e-4)
    rel = lambda a, b: float((a.double().cpu() - b.double()).norm() / b.double().norm().clamp(min=1e-12))
    assert rel(out.d_hidden.float(), h.grad) < 2e-2
    assert rel(out.d_weight, w.grad) < 2e-2


def test_host_resident_hidden_streams_in_and_matches_device_path():
    """policy_step_from_host: hidden states in pinned host memory are copied chunk by chunk ahead of the GEMMs; every
    output must be bit-identical to the step on device-resident hidden states (same kernels, same order)."""
    B, A, H, V, T = 5, 4096, 64, 1024, 0.9  # 20480 tokens = two chunks, the second ragged
    g, hidden, weight, labels, mask = _batch(B, A, H, V, 11)
    cfg = O.make_cfg(policy_loss_type="regular", loss_reduction="token_mean", use_kl_loss=True)
    olp = -torch.rand(B, A, generator=g) * 3
    base = -torch.rand(B, A, generator=g) * 3
    adv = torch.randn(B, A, generator=g)
    common = (weight.to(DEV), labels.to(DEV), olp.to(DEV), adv.to(DEV), cfg)
    kw = dict(loss_mask=mask.to(DEV), base_log_probs=base.to(DEV), temperature=T)
    ref = skyrl_b200.policy_update_step(hidden.to(DEV), *common, **kw)
    pinned = hidden.clone().pin_memory()
    for _ in range(2):  # second call re-uses the cached copy events
        out = skyrl_b200.policy_update_step(pinned, *common, **kw)
        assert out.hidden.is_cuda and torch.equal(out.hidden.cpu().view_as(hidden), hidden)
        assert torch.equal(out.log_probs, ref.log_probs) and torch.equal(out.entropy, ref.entropy)
        assert torch.equal(out.d_hidden, ref.d_hidden) and torch.equal(out.d_weight, ref.d_weight)
        assert out.status == ref.status
    with pytest.raises(ValueError):
        skyrl_b200.policy_update_step(pinned.float(), *common, **kw)  # host path wants bf16 as is


def test_full_size_config2_sampled_rows_and_checksum_identities():
    """BASELINE config 2 at FULL size (256 x 4096 tokens, H = 3584, V = 152064; one fused step, ~3 s):

    * 48 sampled token rows: log-prob / entropy / d_hidden against the CPU oracle evaluated on those rows only;
    * size-independent identities over the whole result (a checksum of checksums):
      sum_v d_weight[v, :] = 0  (softmax - onehot sums to zero over the vocabulary), and
      <d_weight, weight> = <d_hidden, hidden>  (both equal sum_t,v d_logits[t, v] * logits[t, v]);
    * the status dict against the oracle loss evaluated on the kernel's own log-probs.
    """
    free, _ = torch.cuda.mem_get_info()
    if free < 60e9:
        pytest.skip("needs ~40 GB of free HBM")
    B, A, H, V, T = 256, 4096, 3584, 152064, 1.0
    N = B * A
    gen = torch.Generator(device=DEV).manual_seed(2)
    hidden = torch.empty((N, H), dtype=torch.bfloat16, device=DEV)
    for r0 in range(0, N, 65536):
        hidden[r0:r0 + 65536] = torch.randn((min(65536, N - r0), H), generator=gen, device=DEV)
    weight = (torch.randn((V, H), generator=gen, device=DEV) * (1.5 / H ** 0.5)).bfloat16()
    labels = torch.randint(0, V, (B, A), generator=gen, device=DEV)
    lens = torch.randint(A // 4, A + 1, (B,), generator=gen, device=DEV)
    mask = (torch.arange(A, device=DEV)[None] < lens[:, None]).float()
    adv = torch.randn((B, 1), generator=gen, device=DEV).expand(B, A).contiguous()
    cfg = O.make_cfg(policy_loss_type="regular", loss_reduction="token_mean", use_kl_loss=True)
    # old / base log-probs near the current ones: a loss-only pass gives the current log-probs
    lp0 = skyrl_b200.policy_update_step(hidden, weight, labels, torch.zeros(B, A, device=DEV), adv, cfg, loss_mask=mask,
                                        need_d_hidden=False, need_d_weight=False, temperature=T).log_probs
    olp = lp0 + 0.1 * torch.randn((B, A), generator=gen, device=DEV)
    base = lp0 + 0.1 * torch.randn((B, A), generator=gen, device=DEV)
    out = skyrl_b200.policy_update_step(hidden, weight, labels, olp, adv, cfg, loss_mask=mask, base_log_probs=base,
                                        temperature=T)
    assert torch.allclose(out.log_probs, lp0, atol=2e-5, rtol=0)  # the forward does not depend on the loss inputs
    # (last-bit differences only: the forward-only call sums the slab in one online pass, the forward that also
    # stores the tile for the backward takes the slab maximum first)
    # status (1e-5 of the L1 mass), 48 sampled token rows (log-prob, entropy, d_hidden), 16 sampled d_weight rows
    # against the fp32 definition, and the two identities over the whole result
    from step_checker import check_step

    res = check_step(out, hidden, weight, labels.reshape(-1), olp, adv, cfg, mask, base, T, n_sample=48)
    print("full-size config 2:", {k: f"{v:.2e}" for k, v in res.items()})
    assert res["d_weight_rows_rel"] < 1e-2 and res["d_hidden_rel"] < 1e-2 and res["colsum"] < 3e-4


def test_zero_row_skipping_is_bit_identical(tmp_path):
    """The fused step skips 256-row tiles (d_hidden) and 64-row K blocks (d_weight) whose d(logits) rows are all zero
    (masked tokens).  Same batch with the skipping on and off, in two processes: every output bit-identical, and the
    d_hidden rows of masked tokens exactly zero."""
    import os
    import subprocess
    import sys

    worker = os.path.join(os.path.dirname(__file__), "skip_worker.py")
    res = {}
    for flag in ("0", "1"):
        path = str(tmp_path / f"skip{flag}.pt")
        env = dict(os.environ, SKYB200_SKIP_ZERO_ROWS=flag)
        r = subprocess.run([sys.executable, worker, path], env=env, capture_output=True, text=True, timeout=280)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        res[flag] = torch.load(path)
    a, b = res["0"], res["1"]
    for k in ("d_hidden", "d_weight", "d_weight2", "log_probs", "ag_d_hidden", "ag_d_weight"):
        assert torch.equal(a[k], b[k]), k
    assert a["status"] == b["status"]
    dead = a["mask"].reshape(-1) == 0
    assert dead.sum() > 10000 and torch.all(b["d_hidden"][dead] == 0) and torch.all(b["ag_d_hidden"][dead] == 0)
    assert b["d_hidden"][~dead].abs().sum() > 0


def test_wide_backward_tiles_equal_narrow(tmp_path):
    """The 256 x 512 backward tiles (gemm_sm100.cuh, BN == 512: one accumulator in all 512 TMEM columns, two MMA groups
    per K step) against the 256 x 256 ones on a head that is ragged in every tile dimension, in two processes with the
    same chunking: every element accumulates the same K blocks in the same order, so the gradients must agree to fp32
    rounding of the cross-chunk d_weight sums - in the fused step (dense and masked batch, fresh and accumulating
    d_weight) and in the autograd head backward."""
    import os
    import subprocess
    import sys

    worker = os.path.join(os.path.dirname(__file__), "wide_worker.py")
    res = {}
    for flag in ("0", "7"):
        path = str(tmp_path / f"wide{flag}.pt")
        env = dict(os.environ, SKYB200_WIDE=flag, SKYB200_CHUNK_TILES="37")
        r = subprocess.run([sys.executable, worker, path], env=env, capture_output=True, text=True, timeout=280)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        res[flag] = torch.load(path)
    a, b = res["0"], res["7"]
    assert torch.equal(a["log_probs"], b["log_probs"]) and a["status"] == b["status"]
    for k in ("d_hidden", "d_hidden2", "ag_d_hidden", "d_weight", "d_weight2", "ag_d_weight"):
        x, y = a[k].double(), b[k].double()
        assert x.abs().sum() > 0
        assert float((x - y).norm() / x.norm()) < 1e-6, k
