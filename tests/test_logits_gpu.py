"""logprobs_from_logits / chunked_entropy_from_logits on materialised logits (the reference's helper API)."""
import math

import numpy as np
import pytest
import torch

import skyrl_b200
from conftest import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _t(x):
    return torch.from_numpy(np.asarray(x))


def test_match_reference_golden(golden_head):
    g = golden_head
    logits, labels, amask = _t(g["logits"]).to(DEV), _t(g["labels2"]).to(DEV), _t(g["amask"]).to(DEV)
    assert_close(skyrl_b200.logprobs_from_logits(logits.clone(), labels), g["logits_lp"], rel=1e-5)
    assert_close(skyrl_b200.logprobs_from_logits_v2(logits, labels), g["logits_lp"], rel=1e-5)
    ent = skyrl_b200.chunked_entropy_from_logits(logits, attention_mask=amask)
    assert_close(ent, g["logits_ent"], rel=1e-5)
    assert torch.equal(ent.cpu() == 0, _t(g["amask"]) == 0)
    # backward, in place over the logits buffer like inplace_backward=True
    z = logits.clone().requires_grad_(True)
    zz = z * 1.0
    buf_ptr = zz.data_ptr()
    (skyrl_b200.logprobs_from_logits(zz, labels, inplace_backward=True) * _t(g["u"]).to(DEV)).sum().backward()
    assert_close(z.grad, g["dlogits_lp"], rel=1e-5, floor=0.1)
    assert buf_ptr == zz.data_ptr()
    z2 = logits.clone().requires_grad_(True)
    (skyrl_b200.chunked_entropy_from_logits(z2, requires_grad=True, attention_mask=amask) * _t(g["v"]).to(DEV)).sum().backward()
    assert_close(z2.grad, g["dlogits_ent"], rel=1e-5, floor=0.1)
    assert torch.equal(logits.cpu(), _t(g["logits"])), "non-inplace paths must not touch the logits"


@pytest.mark.parametrize("entropy_first", [False, True])
def test_inplace_logprob_backward_does_not_corrupt_the_entropy_gradient(entropy_first):
    """HFModelWrapper.forward calls logprobs_from_logits(logits, inplace_backward=True) AND
    chunked_entropy_from_logits(logits, requires_grad=use_entropy_loss) on the SAME buffer (model_wrapper.py:317, 347).
    Whichever backward autograd runs first, both gradients must be those of the untouched logits: the log-prob
    backward only overwrites the buffer once no entropy node still needs it."""
    g = torch.Generator(device=DEV).manual_seed(3)
    B, S, V = 2, 7, 1000
    z0 = torch.randn(B, S, V, device=DEV, generator=g) * 2.0
    lab = torch.randint(0, V, (B, S), device=DEV, generator=g)
    am = (torch.rand(B, S, device=DEV, generator=g) > 0.2).float()
    u, v = torch.randn(B, S, device=DEV, generator=g), torch.randn(B, S, device=DEV, generator=g)
    zr = z0.clone().requires_grad_(True)
    lsm = torch.log_softmax(zr, -1)
    ref = (lsm.gather(-1, lab[..., None]).squeeze(-1) * u).sum() + ((-(lsm.exp() * lsm).sum(-1) * am) * v).sum()
    ref.backward()
    z = z0.clone().requires_grad_(True)
    logits = z * 1.0  # the lm_head output: a non-leaf buffer the in-place backward may overwrite
    if entropy_first:
        ent = skyrl_b200.chunked_entropy_from_logits(logits, requires_grad=True, attention_mask=am)
        lp = skyrl_b200.logprobs_from_logits(logits, lab, inplace_backward=True)
    else:
        lp = skyrl_b200.logprobs_from_logits(logits, lab, inplace_backward=True)
        ent = skyrl_b200.chunked_entropy_from_logits(logits, requires_grad=True, attention_mask=am)
    ((lp * u).sum() + (ent * v).sum()).backward()
    assert_close(z.grad, zr.grad, rel=1e-5, floor=0.1)
    from skyrl_b200 import torch_utils

    assert not torch_utils._PENDING_ENTROPY  # nothing left registered
    # without a grad-requiring entropy on the buffer the backward IS in place (the reference's memory behaviour)
    z = z0.clone().requires_grad_(True)
    logits = z * 1.0
    before = logits._version
    (skyrl_b200.logprobs_from_logits(logits, lab, inplace_backward=True) * u).sum().backward()
    assert logits._version > before  # ... and autograd can see it: the version counter moved


def test_cross_entropy_loss_accepts_none_like_the_reference():
    """ppo_utils.py:892-930 documents old_log_probs / advantages as ignored for SFT; callers pass None."""
    from types import SimpleNamespace as NS

    g = torch.Generator(device=DEV).manual_seed(4)
    lp = (-torch.rand(3, 11, device=DEV, generator=g)).requires_grad_(True)
    mask = (torch.rand(3, 11, device=DEV, generator=g) > 0.3).float()
    loss, met = skyrl_b200.PolicyLossRegistry.get("cross_entropy")(lp, None, None, NS(loss_reduction="token_mean"),
                                                                   loss_mask=mask)
    loss.backward()
    assert loss.item() == pytest.approx(float((-lp.detach() * mask).sum()), rel=1e-5)
    assert torch.allclose(lp.grad, -mask) and met == {"clip_ratio": 0.0}


def test_uniform_logits_known_answer():  # reference: tests/cpu/utils/test_torch_utils.py:33-55
    ent = skyrl_b200.chunked_entropy_from_logits(torch.zeros(2, 3, 4, device=DEV))
    assert torch.allclose(ent.cpu(), torch.full((2, 3), math.log(4.0)), atol=1e-4)
    with pytest.raises(ValueError, match="does not match logits shape"):
        skyrl_b200.chunked_entropy_from_logits(torch.zeros(2, 3, 4, device=DEV),
                                               attention_mask=torch.ones(2, 4, device=DEV))


def test_bf16_logits_and_output_dtype():
    g = torch.Generator(device=DEV).manual_seed(0)
    z = (torch.randn(3, 5, 151936, device=DEV, generator=g) * 1.5).bfloat16()
    lab = torch.randint(0, 151936, (3, 5), device=DEV, generator=g)
    lp = skyrl_b200.logprobs_from_logits(z.clone(), lab)
    ref = torch.log_softmax(z.float(), -1).gather(-1, lab[..., None]).squeeze(-1)
    assert lp.dtype == torch.float32 and (lp - ref).abs().max().item() < 1e-4
    ent = skyrl_b200.chunked_entropy_from_logits(z)
    assert ent.dtype == torch.bfloat16  # the reference stores entropy in the logits dtype (torch_utils.py:92)
    p = torch.softmax(z.float(), -1)
    ref_e = -(p * torch.log_softmax(z.float(), -1)).sum(-1)
    assert (ent.float() - ref_e).abs().max().item() < 0.05


def test_masked_mean_kernel():
    g = torch.Generator(device=DEV).manual_seed(1)
    x = torch.randn(37, 1001, device=DEV, generator=g)
    m = (torch.rand(37, 1001, device=DEV, generator=g) > 0.5).float()
    assert_close(skyrl_b200.masked_mean(x, m), (x * m).sum() / m.sum(), rel=1e-5, floor=1e-3)
    assert skyrl_b200.masked_mean(x, torch.zeros_like(m)).item() == 0.0  # denominator clamped to 1
    assert_close(skyrl_b200.masked_mean(x, m, dim=-1), (x * m).sum(-1) / m.sum(-1).clamp(min=1), rel=1e-5, floor=1e-3)


def test_value_head_forward_backward():
    from skyrl_b200.value_head import action_values, value_head

    g = torch.Generator(device=DEV).manual_seed(5)
    B, S, H, A = 3, 50, 200, 17
    hidden = torch.randn(B, S, H, device=DEV, generator=g).bfloat16().requires_grad_(True)
    lin = torch.nn.Linear(H, 1).to(DEV)
    v = value_head(hidden, lin.weight, lin.bias)
    ref = torch.nn.functional.linear(hidden.float(), lin.weight, lin.bias).squeeze(-1)
    assert_close(v, ref, rel=1e-5, floor=1.0)
    gv = torch.randn(B, S, device=DEV, generator=g)
    (v * gv).sum().backward()
    dh, dw, db = hidden.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone()
    hidden.grad = None
    lin.zero_grad()
    h32 = hidden.detach().float().requires_grad_(True)
    (torch.nn.functional.linear(h32, lin.weight, lin.bias).squeeze(-1) * gv).sum().backward()
    assert_close(dh.float(), h32.grad, rel=1e-2, floor=1e-2)  # bf16 output
    assert_close(dw, lin.weight.grad, rel=1e-4, floor=1e-2)
    assert_close(db, lin.bias.grad, rel=1e-5, floor=1e-2)
    av = action_values(hidden.detach(), lin.weight, A, lin.bias)
    assert torch.equal(av, v.detach()[:, :-1][:, -A:])
