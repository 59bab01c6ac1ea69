"""Runs one fused policy step on a heavily padded batch and saves the results (used by
test_zero_row_skipping_is_bit_identical with SKYB200_SKIP_ZERO_ROWS=0 / 1 in two processes)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import skyrl_b200  # noqa: E402
from oracle import policy_oracle as O  # noqa: E402


def main(out_path):
    dev = "cuda:0"
    B, A, H, V = 6, 4096, 64, 1024   # 24576 tokens = 3 chunks (the last ragged)
    g = torch.Generator().manual_seed(4)
    hidden = torch.randn(B, A, H, generator=g).bfloat16()
    weight = (torch.randn(V, H, generator=g) * (1.5 / H ** 0.5)).bfloat16()
    labels = torch.randint(0, V, (B, A), generator=g)
    lens = torch.tensor([4096, 100, 0, 1300, 257, 2048])      # long masked tails, one empty sequence
    mask = (torch.arange(A)[None] < lens[:, None]).float()
    mask[0, 1000:1700] = 0.0                                   # an interior hole spanning whole tiles
    olp = -torch.rand(B, A, generator=g) * 3
    adv = torch.randn(B, A, generator=g)
    cfg = O.make_cfg(policy_loss_type="regular", loss_reduction="sequence_mean", use_kl_loss=True,
                     use_entropy_loss=True, entropy_loss_coef=0.02)
    dw = torch.full((V, H), 0.5, device=dev)
    out = skyrl_b200.policy_update_step(hidden.to(dev), weight.to(dev), labels.to(dev), olp.to(dev), adv.to(dev), cfg,
                                        loss_mask=mask.to(dev), base_log_probs=olp.to(dev), d_weight=dw,
                                        accumulate_d_weight=True)
    out2 = skyrl_b200.policy_update_step(hidden.to(dev), weight.to(dev), labels.to(dev), olp.to(dev), adv.to(dev), cfg,
                                         loss_mask=mask.to(dev), base_log_probs=olp.to(dev))   # fresh d_weight
    # the autograd path (fused_lm_head + a masked loss): skyb200_head_backward skips the same rows in its 3 GEMMs
    h = hidden.to(dev).reshape(B * A, H).requires_grad_(True)
    w = weight.to(dev).requires_grad_(True)
    lp, ent = skyrl_b200.fused_lm_head(h, w, labels.to(dev).reshape(-1), temperature=0.9)
    m = mask.to(dev).reshape(-1)
    ((lp * adv.to(dev).reshape(-1) * m).sum() / m.sum() - 0.01 * (ent * m).sum() / m.sum()).backward()
    torch.save({"ag_d_hidden": h.grad.cpu(), "ag_d_weight": w.grad.cpu(), "d_hidden": out.d_hidden.cpu(), "d_weight": out.d_weight.cpu(), "d_weight2": out2.d_weight.cpu(),
                "log_probs": out.log_probs.cpu(), "status": out.status, "mask": mask}, out_path)


if __name__ == "__main__":
    main(sys.argv[1])
