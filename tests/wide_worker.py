"""Runs one fused policy step and one autograd head backward on a batch whose head shape is ragged in every tile
dimension, and saves the results (used by test_wide_backward_tiles_equal_narrow with SKYB200_WIDE=0 / 7 in two processes)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import skyrl_b200  # noqa: E402
from oracle import policy_oracle as O  # noqa: E402


def main(out_path):
    dev = "cuda:0"
    B, A, H, V = 5, 2000, 1064, 3000   # 10000 tokens = 2 chunks of 37 pair-tiles; H = 2 wide tiles + 40 columns
    g = torch.Generator().manual_seed(11)
    hidden = torch.randn(B, A, H, generator=g).bfloat16()
    weight = (torch.randn(V, H, generator=g) * (1.5 / H ** 0.5)).bfloat16()
    labels = torch.randint(0, V, (B, A), generator=g)
    lens = torch.tensor([2000, 700, 1999, 0, 1300])
    mask = (torch.arange(A)[None] < lens[:, None]).float()
    olp = -torch.rand(B, A, generator=g) * 3
    adv = torch.randn(B, A, generator=g)
    cfg = O.make_cfg(policy_loss_type="regular", loss_reduction="token_mean", use_kl_loss=True,
                     use_entropy_loss=True, entropy_loss_coef=0.02)
    dw = torch.full((V, H), 0.25, device=dev)
    args = (hidden.to(dev), weight.to(dev), labels.to(dev), olp.to(dev), adv.to(dev), cfg)
    out = skyrl_b200.policy_update_step(*args, loss_mask=mask.to(dev), base_log_probs=olp.to(dev), d_weight=dw,
                                        accumulate_d_weight=True)
    out2 = skyrl_b200.policy_update_step(*args, loss_mask=torch.ones(B, A, device=dev), base_log_probs=olp.to(dev))
    h = hidden.to(dev).reshape(B * A, H).requires_grad_(True)
    w = weight.to(dev).requires_grad_(True)
    lp, ent = skyrl_b200.fused_lm_head(h, w, labels.to(dev).reshape(-1), temperature=0.9)
    ((lp * adv.to(dev).reshape(-1)).mean() - 0.01 * ent.mean()).backward()
    torch.save({"ag_d_hidden": h.grad.cpu(), "ag_d_weight": w.grad.cpu(), "d_hidden": out.d_hidden.cpu(),
                "d_weight": out.d_weight.cpu(), "d_hidden2": out2.d_hidden.cpu(), "d_weight2": out2.d_weight.cpu(),
                "log_probs": out.log_probs.cpu(), "status": out.status}, out_path)


if __name__ == "__main__":
    main(sys.argv[1])
