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
