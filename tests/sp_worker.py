"""Worker of tests/test_dist_cpu.py::test_sequence_parallel_wrapper_glue (gloo, CPU): the Ulysses branch of the wrapper
glue with the CUDA head replaced by a differentiable stand-in, against the single-process padded-layout glue."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def stand_in_head(h, w, labels, temperature=1.0, entropy_mask=None, compute_entropy=True, chunk=0):
    """Per-token, differentiable in h and w, label- and temperature-dependent (like the real head)."""
    z = (h.float() @ w.float().t()) / temperature
    lp = z.gather(1, labels.view(-1, 1) % w.shape[0]).squeeze(1) - torch.logsumexp(z, dim=1)
    ent = z.pow(2).mean(1)
    if entropy_mask is not None:
        ent = ent * entropy_mask.float()
    return lp, ent


def main():
    rank, world, port, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", port
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from skyrl_b200 import model_wrapper as mw

        mw.fused_lm_head = stand_in_head
        res = {}
        for case, (B, S, A, Hd, V) in enumerate([(3, 13, 4, 8, 11), (2, 16, 3, 8, 11), (2, 9, 7, 4, 5)]):
            g = torch.Generator().manual_seed(100 + case)           # same data on every rank
            hidden = torch.randn(B, S, Hd, generator=g)
            weight = torch.randn(V, Hd, generator=g)
            sequences = torch.randint(0, V, (B, S), generator=g)
            am = (torch.rand(B, S, generator=g) > 0.2).long()
            coef = torch.randn(B, A, generator=g)
            # single-process reference: the padded-layout glue on the full hidden states
            h0 = hidden.clone().requires_grad_(True)
            w0 = weight.clone().requires_grad_(True)
            lp0, ent0 = mw.action_log_probs_from_hidden(h0, w0, sequences, A, attention_mask=am, temperature=0.7,
                                                        compute_entropy=True)
            ((lp0 * coef).sum() + 0.3 * (ent0 * coef).sum()).backward()
            # this rank's slice of the (zero-padded) sequence dimension
            pad = (world - S % world) % world
            L = (S + pad) // world
            hp = torch.nn.functional.pad(hidden, (0, 0, 0, pad))
            h1 = hp[:, rank * L:(rank + 1) * L].clone().requires_grad_(True)
            w1 = weight.clone().requires_grad_(True)
            lp1, ent1 = mw.action_log_probs_from_sp_hidden(h1, w1, sequences, A, attention_mask=am, temperature=0.7,
                                                           compute_entropy=True, grad_scaler=(case != 1))
            ((lp1 * coef).sum() + 0.3 * (ent1 * coef).sum()).backward()
            scale = 1.0 if case == 1 else float(world)
            want_h = torch.nn.functional.pad(h0.grad, (0, 0, 0, pad))[:, rank * L:(rank + 1) * L] * scale
            dw = w1.grad.clone() if w1.grad is not None else torch.zeros_like(weight)
            dist.all_reduce(dw)                                       # the ranks' weight gradients add up
            res[case] = dict(
                fwd=bool(torch.allclose(lp1, lp0, atol=1e-6) and torch.allclose(ent1, ent0, atol=1e-6)),
                dh=bool(h1.grad is not None and torch.allclose(h1.grad, want_h, atol=1e-5)),
                dw=bool(torch.allclose(dw, w0.grad * scale, atol=1e-4)),
                rows_here=int(min(max(S - 1 - rank * L, 0), L) - min(max(S - A - 1 - rank * L, 0), L)))
        # ---- sample packing + sequence parallelism against the single-process packed glue
        for case, (B, S, A, Hd, V) in enumerate([(3, 13, 4, 8, 11), (2, 16, 9, 8, 11), (4, 9, 2, 4, 5)]):
            g = torch.Generator().manual_seed(200 + case)
            weight = torch.randn(V, Hd, generator=g)
            sequences = torch.randint(0, V, (B, S), generator=g)
            am = (torch.rand(B, S, generator=g) > 0.25).long()
            am[0] = 1
            nnz = int(am.sum())
            packed = torch.randn(nnz, Hd, generator=g)
            coef = torch.randn(B, A, generator=g)
            h0 = packed.clone().requires_grad_(True)
            w0 = weight.clone().requires_grad_(True)
            lp0, ent0 = mw.action_log_probs_from_packed_hidden(h0, w0, sequences, am, A, temperature=0.7,
                                                               compute_entropy=True)
            ((lp0 * coef).sum() + 0.3 * (ent0 * coef).sum()).backward()
            pad = (world - nnz % world) % world
            L = (nnz + pad) // world
            hp = torch.nn.functional.pad(packed, (0, 0, 0, pad))
            h1 = hp[rank * L:(rank + 1) * L].clone().requires_grad_(True)
            w1 = weight.clone().requires_grad_(True)
            lp1, ent1 = mw.action_log_probs_from_sp_packed_hidden(h1, w1, sequences, am, A, temperature=0.7,
                                                                  compute_entropy=True)
            ((lp1 * coef).sum() + 0.3 * (ent1 * coef).sum()).backward()
            want_h = torch.nn.functional.pad(h0.grad, (0, 0, 0, pad))[rank * L:(rank + 1) * L] * world
            dw = w1.grad.clone() if w1.grad is not None else torch.zeros_like(weight)
            dist.all_reduce(dw)
            res[10 + case] = dict(
                fwd=bool(torch.allclose(lp1, lp0, atol=1e-6) and torch.allclose(ent1, ent0, atol=1e-6)),
                dh=bool(h1.grad is not None and torch.allclose(h1.grad, want_h, atol=1e-5)),
                dw=bool(torch.allclose(dw, w0.grad * world, atol=1e-4)), rows_here=1)
        torch.save(res, f"{out}.{rank}")
    finally:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
