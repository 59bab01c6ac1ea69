"""2-rank NCCL worker for test_fused_dweight_reduce_scatter_over_nccl_two_ranks (launched with torchrun): the shards
the d_weight GEMM epilogue reduce-scatters over NVLink equal NCCL's mean reduce-scatter of the per-rank local
gradients (FSDP semantics, fsdp_strategy.py:151-153, 214-226) - one chunk, several chunks, masked rows (zero-row
skipping), two micro-batches accumulated, two optimizer steps (begin() re-zeroes)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import skyrl_b200  # noqa: E402
from skyrl_b200 import dist as sdist  # noqa: E402
from oracle import policy_oracle as O  # noqa: E402  (config namespace only)


REMOTE = os.environ.get("SKYB200_TEST_DWS_REMOTE", "tma")


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    failed = []

    def chk(name, cond):
        if not bool(cond):
            failed.append(name)
            print(f"rank {rank}: check failed: {name}", flush=True)

    rel = lambda x, y: float((x.double() - y.double()).norm() / y.double().norm().clamp(min=1e-30))
    cfg = O.make_cfg(policy_loss_type="regular", loss_reduction="token_mean", use_kl_loss=False)
    for case, (B, A, H, V, ragged) in enumerate([(4, 300, 128, 2048, False), (6, 4000, 256, 4096, True),
                                                 (3, 10000, 64, 64 * world * 5, True),
                                                 # H = 1024, 64 row tiles: the d_weight GEMM takes the wide 256 x 512 tiles
                                                 (2, 3000, 1024, 16384, True)]):
        g = torch.Generator().manual_seed(100 + 7 * case + rank)      # different data on every rank
        hidden = torch.randn(B, A, H, generator=g).bfloat16().to(dev)
        gw = torch.Generator().manual_seed(55 + case)                 # the same weight on every rank
        weight = (torch.randn(V, H, generator=gw) * (1.5 / H ** 0.5)).bfloat16().to(dev)
        labels = torch.randint(0, V, (B, A), generator=g).to(dev)
        adv = torch.randn(B, 1, generator=g).expand(B, A).contiguous().to(dev)
        olp = (-6.0 + 0.3 * torch.randn(B, A, generator=g)).to(dev)
        mask = torch.ones(B, A)
        if ragged:
            lens = torch.randint(A // 5, A + 1, (B,), generator=g)
            mask = (torch.arange(A)[None] < lens[:, None]).float()
        mask = mask.to(dev)
        # reference result: local fp32 gradient, then NCCL's mean reduce-scatter
        loc = skyrl_b200.policy_update_step(hidden, weight, labels, olp, adv, cfg, loss_mask=mask)
        want, _ = sdist.reduce_scatter_d_weight(loc.d_weight.clone())
        sc = sdist.DWeightScatter(V, H, remote=REMOTE)
        chk(f"case {case}: shard shape", tuple(sc.shard.shape) == (V // world, H))
        for opt_step in range(2):                                     # begin() must re-zero between optimizer steps
            sc.begin()
            out = skyrl_b200.policy_update_step(hidden, weight, labels, olp, adv, cfg, loss_mask=mask, dw_scatter=sc,
                                                sync=False)
            shard = sc.finish()
            chk(f"case {case} step {opt_step}: no local d_weight", out.d_weight is None)
            e = rel(shard, want)
            chk(f"case {case} step {opt_step}: fused shard vs NCCL reduce-scatter, rel L2 {e:.3e}", e < 1e-5)
            chk(f"case {case} step {opt_step}: d_hidden unchanged", torch.equal(out.d_hidden, loc.d_hidden))
            chk(f"case {case} step {opt_step}: log-probs unchanged", torch.equal(out.log_probs, loc.log_probs))
        # two micro-batches between begin() and finish(): contributions accumulate
        sc.begin()
        for _ in range(2):
            skyrl_b200.policy_update_step(hidden, weight, labels, olp, adv, cfg, loss_mask=mask, dw_scatter=sc, sync=False)
        e = rel(sc.finish(), 2.0 * want)
        chk(f"case {case}: two accumulated micro-batches, rel L2 {e:.3e}", e < 1e-5)
        sc.close()
    flag = torch.tensor([0.0 if failed else 1.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DWS_OK" if flag.item() == 1.0 else "DWS_FAIL")
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
