"""2-rank NCCL worker for test_data_parallel_step_over_nccl_two_ranks (launched with torchrun): the N-GPU result of the
policy update equals the 1-GPU result on the same batch (SURVEY section 4 take-away, section 8e)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import skyrl_b200  # noqa: E402
from skyrl_b200 import dist as sdist  # noqa: E402
from oracle import policy_oracle as O  # noqa: E402  (config namespace only)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    B, A, H, V, G = 16, 300, 128, 3000, 4
    g = torch.Generator().manual_seed(5)                      # the same global batch on every rank
    hidden = torch.randn(B, A, H, generator=g).bfloat16().to(dev)
    weight = (torch.randn(V, H, generator=g) * (1.5 / H ** 0.5)).bfloat16().to(dev)
    labels = torch.randint(0, V, (B, A), generator=g).to(dev)
    uids = np.array([f"p{i // G}" for i in range(B)])[torch.randperm(B, generator=g).numpy()]
    mask = torch.ones(B, A, device=dev)                        # dense: every shard has the same token count, so the
    rewards = torch.zeros(B, A)                                # mean of per-rank token means is the global token mean
    rewards[:, -1] = torch.randn(B, generator=g)
    rewards = rewards.to(dev)
    values = torch.randn(B, A, generator=g).to(dev)
    olp = (-6.0 + 0.3 * torch.randn(B, A, generator=g)).to(dev)
    cfg = O.make_cfg(policy_loss_type="regular", loss_reduction="token_mean", use_kl_loss=False)
    ok = True

    # ---- 1 GPU: the whole batch
    adv_all, _ = skyrl_b200.compute_grpo_outcome_advantage(rewards, mask, uids)
    whole = skyrl_b200.policy_update_step(hidden, weight, labels, olp, adv_all, cfg, loss_mask=mask)
    gae_all, ret_all = skyrl_b200.compute_gae_advantage_return(rewards, values, mask, 1.0, 0.95)

    # ---- N GPUs: whole prompt groups per rank, nothing exchanged for the advantages
    rows = torch.tensor(sdist.shard_prompt_groups(uids, [A] * B, world)[rank], device=dev)
    adv, _ = skyrl_b200.compute_grpo_outcome_advantage(rewards[rows], mask[rows], uids[rows.cpu().numpy()])
    ok = ok and torch.equal(adv, adv_all[rows])
    part = skyrl_b200.policy_update_step(hidden[rows], weight, labels[rows], olp[rows], adv, cfg, loss_mask=mask[rows],
                                         sync=False)
    stats = sdist.all_reduce_packed_stats(part.stats)          # mean of per-rank means
    sdist.all_reduce_d_weight(part.d_weight)                   # mean of per-rank gradients (in place)
    from skyrl_b200 import _lib

    ok = ok and abs(stats[_lib.STAT_FINAL_LOSS].item() - whole.status["final_loss"]) < 1e-5 * max(1.0, abs(whole.status["final_loss"]))
    rel = lambda x, y: float((x.float() - y.float()).norm() / y.float().norm())
    ok = ok and rel(part.d_weight, whole.d_weight) < 2e-3      # bf16 d_logits are rounded after different 1/n scales
    ok = ok and rel(part.d_hidden.float() / world, whole.d_hidden[rows].float()) < 2e-3
    ok = ok and torch.equal(part.log_probs, whole.log_probs[rows])

    # ---- GAE: rows are local, the whitening statistics are global (3 moments all-reduced)
    a_loc, r_loc = sdist.gae_advantage_sharded(rewards[rows], values[rows], mask[rows], 1.0, 0.95)
    ok = ok and torch.equal(r_loc, ret_all[rows]) and (a_loc - gae_all[rows]).abs().max().item() < 1e-5

    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DP_OK" if flag.item() == 1.0 else "DP_FAIL")
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
