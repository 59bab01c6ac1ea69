"""2-rank NCCL worker of test_sequence_parallel_glue_over_nccl_two_ranks (launched with torchrun): the Ulysses branch
of the wrapper glue (model_wrapper.py:292-355 of the reference) with the REAL fused head on two GPUs against the
single-GPU padded glue on the same data - the reference's own SP=2 vs SP=1 check
(tests/gpu/gpu_ci/test_model_wrapper.py:134-191).  The gloo twin with a stand-in head is tests/sp_worker.py."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from skyrl_b200 import model_wrapper as mw  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    rel = lambda x, y: float((x.float() - y.float()).norm() / y.float().norm().clamp(min=1e-30))
    # (B, S, A, H, V): a window inside one slice, a window that straddles both slices (S needs padding), A = S - 1
    for case, (B, S, A, Hd, V) in enumerate([(3, 301, 40, 128, 3000), (2, 257, 200, 64, 1000), (2, 64, 63, 128, 2000)]):
        g = torch.Generator().manual_seed(300 + case)           # same data on every rank
        hidden = torch.randn(B, S, Hd, generator=g).bfloat16().to(dev)
        weight = (torch.randn(V, Hd, generator=g) * (1.5 / Hd ** 0.5)).bfloat16().to(dev)
        sequences = torch.randint(0, V, (B, S), generator=g).to(dev)
        am = (torch.rand(B, S, generator=g) > 0.1).long().to(dev)
        coef = torch.randn(B, A, generator=g).to(dev)
        h0 = hidden.clone().requires_grad_(True)
        w0 = weight.clone().requires_grad_(True)
        lp0, ent0 = mw.action_log_probs_from_hidden(h0, w0, sequences, A, attention_mask=am, temperature=0.7,
                                                    compute_entropy=True)
        ((lp0 * coef).sum() + 0.3 * (ent0 * coef).sum()).backward()
        pad = (world - S % world) % world
        L = (S + pad) // world
        hp = torch.nn.functional.pad(hidden, (0, 0, 0, pad))
        h1 = hp[:, rank * L:(rank + 1) * L].clone().requires_grad_(True)
        w1 = weight.clone().requires_grad_(True)
        lp1, ent1 = mw.action_log_probs_from_sp_hidden(h1, w1, sequences, A, attention_mask=am, temperature=0.7,
                                                       compute_entropy=True, grad_scaler=True)
        ((lp1 * coef).sum() + 0.3 * (ent1 * coef).sum()).backward()
        # forward: the same per-token kernel on the same rows -> bit-identical after the all-gather
        ok = ok and torch.equal(lp1, lp0.detach()) and torch.equal(ent1, ent0.detach())
        want_h = torch.nn.functional.pad(h0.grad.float(), (0, 0, 0, pad))[:, rank * L:(rank + 1) * L] * world
        ok = ok and h1.grad is not None and rel(h1.grad, want_h) < 5e-3
        dw = (w1.grad.float() if w1.grad is not None else torch.zeros_like(weight, dtype=torch.float32)).clone()
        dist.all_reduce(dw)                                       # the ranks' weight gradients add up
        ok = ok and rel(dw, w0.grad.float() * world) < 5e-3
        if not ok:
            print(f"rank {rank} case {case} FAILED", flush=True)
            break
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SP_OK" if flag.item() == 1.0 else "SP_FAIL")
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
