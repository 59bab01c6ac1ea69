"""2-rank NCCL worker for test_vocab_parallel_head_over_nccl_two_ranks (launched with torchrun)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import skyrl_b200  # noqa: E402
from skyrl_b200 import tp_head  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    N, H, V, T = 600, 128, 4096, 0.9
    g = torch.Generator().manual_seed(3)  # same data on every rank (hidden is replicated in a TP group)
    hidden = torch.randn(N, H, generator=g).bfloat16().to(dev)
    weight = (torch.randn(V, H, generator=g) * (1.5 / H ** 0.5)).bfloat16().to(dev)
    labels = torch.randint(0, V, (N,), generator=g).to(dev)
    g_lp = torch.randn(N, generator=g).to(dev)
    hf = hidden.clone().requires_grad_(True)
    wf = weight.clone().requires_grad_(True)
    lp_full, ent_full = skyrl_b200.fused_lm_head(hf, wf, labels, temperature=T)
    (lp_full * g_lp).sum().backward()
    per = V // world
    a, b = rank * per, (rank + 1) * per
    hs = hidden.clone().requires_grad_(True)
    ws = weight[a:b].clone().requires_grad_(True)
    lp, ent = tp_head.vocab_parallel_lm_head(hs, ws, labels, a, temperature=T)
    (lp * g_lp).sum().backward()
    ok = (lp - lp_full.detach()).abs().max().item() < 1e-5 and (ent - ent_full.detach()).abs().max().item() < 1e-5
    rel = lambda x, y: float((x.float() - y.float()).norm() / y.float().norm())
    ok = ok and rel(hs.grad, hf.grad) < 1e-2 and rel(ws.grad, wf.grad[a:b]) < 1e-2
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("TP_OK" if flag.item() == 1.0 else "TP_FAIL")
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
