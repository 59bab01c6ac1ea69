"""The large-shape step checker (tests/step_checker.py) checked on the CPU: an oracle-built step result must pass it
in both layouts, and the corruptions it exists to catch must fail it."""
from types import SimpleNamespace

import pytest
import torch

from oracle import policy_oracle as O
from step_checker import check_step


def _oracle_step(kw, packed, seed=0):
    B, A, H, V, T = 4, 48, 32, 300, 0.8
    g = torch.Generator().manual_seed(seed)
    hidden = torch.randn(B, A, H, generator=g).bfloat16()
    weight = (torch.randn(V, H, generator=g) * (1.5 / H ** 0.5)).bfloat16()
    labels = torch.randint(0, V, (B, A), generator=g)
    lens = torch.tensor([48, 20, 1, 33])
    keep = torch.arange(A)[None] < lens[:, None]
    mask = keep.float()
    mask[0, 5:9] = 0
    cfg = O.make_cfg(**kw)
    h = hidden.float().reshape(B * A, H).requires_grad_(True)
    w = weight.float().requires_grad_(True)
    lp, ent = O.head_logprobs_entropy(h, w, labels.reshape(-1), T)
    lp, ent = lp.view(B, A), ent.view(B, A)
    olp = lp.detach() + 0.05 * torch.randn(B, A, generator=g)
    base = lp.detach() + 0.1 * torch.randn(B, A, generator=g)
    adv = torch.randn(B, A, generator=g)
    final, st = O.loss_assembly(lp, olp, adv, cfg, mask, base, ent)
    final.backward()
    sel = keep.reshape(-1) if packed else torch.ones(B * A, dtype=torch.bool)
    shape = (-1,) if packed else (B, A)
    out = SimpleNamespace(status=st, log_probs=lp.detach().reshape(-1)[sel].reshape(shape),
                          entropy=ent.detach().reshape(-1)[sel].reshape(shape),
                          d_hidden=h.grad[sel].clone(), d_weight=w.grad.clone())
    args = (hidden.reshape(B * A, H)[sel], weight, labels.reshape(-1)[sel], olp, adv, cfg, mask, base, T)
    return out, args, (keep if packed else None)


KW = [dict(policy_loss_type="regular", loss_reduction="token_mean", use_kl_loss=True, use_entropy_loss=True),
      dict(policy_loss_type="gspo", loss_reduction="sequence_mean", use_kl_loss=False),
      dict(policy_loss_type="dual_clip", loss_reduction="seq_mean_token_sum_norm", max_seq_len=64, use_kl_loss=True)]


@pytest.mark.parametrize("kw", KW)
@pytest.mark.parametrize("packed", [False, True])
def test_checker_accepts_the_oracle_result(kw, packed):
    if packed and kw["policy_loss_type"] == "gspo":
        pytest.skip("the packed layout has no gspo")
    out, args, keep = _oracle_step(kw, packed)
    res = check_step(out, *args, keep=keep, n_sample=16)
    assert res["lp_err"] < 1e-5 and res["d_hidden_rel"] < 1e-5 and res["inner_err"] < 1e-6


@pytest.mark.parametrize("what", ["lost_dw_tile", "lost_dh_rows", "wrong_loss", "shifted_logprob", "no_entropy_grad"])
def test_checker_rejects_corruptions(what):
    out, args, keep = _oracle_step(KW[0], packed=False)
    if what == "lost_dw_tile":
        out.d_weight[64:128, :16] = 0
    elif what == "lost_dh_rows":
        out.d_hidden[:] = out.d_hidden * 0.9
    elif what == "wrong_loss":
        out.status = dict(out.status, final_loss=out.status["final_loss"] + 1e-3)
    elif what == "shifted_logprob":
        out.log_probs = out.log_probs + 5e-3
    elif what == "no_entropy_grad":  # d_hidden of a step that forgot the entropy-bonus gradient
        out2, _, _ = _oracle_step(dict(KW[0], use_entropy_loss=False), packed=False)
        out.d_hidden = out2.d_hidden
    with pytest.raises(AssertionError):
        check_step(out, *args, keep=keep, n_sample=16, grad_rel=2e-3 if what == "no_entropy_grad" else 2e-2)
