"""One fused policy step on sharply peaked logits, run once per value of SKYB200_STORED_LSE (the flag is read once per
process); saves what tests/test_zconfigs_gpu.py::test_stored_lse_flag compares."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import skyrl_b200  # noqa: E402
from oracle import policy_oracle as O  # noqa: E402


def main(out_path):
    dev = "cuda:0"
    res = {}
    for name, std, T in (("peaked", 8.0, 0.6), ("mild", 1.5, 1.0)):
        B, A, H, V = 3, 400, 256, 5000   # 1200 tokens: one chunk with a ragged last tile; V is not a slab multiple
        g = torch.Generator().manual_seed(9)
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
        out = skyrl_b200.policy_update_step(hidden.to(dev), weight.to(dev), labels.to(dev).view(B, A), olp.to(dev),
                                            adv.to(dev), cfg, loss_mask=mask.to(dev), temperature=T)
        rel = lambda a, b: float((a.double().cpu() - b.double()).norm() / b.double().norm())
        dw = out.d_weight.double().cpu()
        res[name] = dict(log_probs=out.log_probs.cpu(), entropy=out.entropy.cpu(), final=out.status["final_loss"],
                         final_ref=final.item(), lp_err=float((out.log_probs.cpu().view(-1) - lp_ref.detach()).abs().max()),
                         rel_dh=rel(out.d_hidden.float(), h.grad), rel_dw=rel(out.d_weight, w.grad),
                         colsum=float((dw.sum(0).abs() / dw.abs().sum(0)).max()))
    torch.save(res, out_path)


if __name__ == "__main__":
    main(sys.argv[1])
