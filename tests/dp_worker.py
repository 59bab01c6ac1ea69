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
    failed = []

    def chk(name, cond):
        if not bool(cond):
            failed.append(name)
            print(f"rank {rank}: check failed: {name}", flush=True)

    # ---- 1 GPU: the whole batch
    adv_all, _ = skyrl_b200.compute_grpo_outcome_advantage(rewards, mask, uids)
    whole = skyrl_b200.policy_update_step(hidden, weight, labels, olp, adv_all, cfg, loss_mask=mask)
    gae_all, ret_all = skyrl_b200.compute_gae_advantage_return(rewards, values, mask, 1.0, 0.95)

    # ---- N GPUs: whole prompt groups per rank, nothing exchanged for the advantages
    rows = torch.tensor(sdist.shard_prompt_groups(uids, [A] * B, world)[rank], device=dev)
    adv, _ = skyrl_b200.compute_grpo_outcome_advantage(rewards[rows], mask[rows], uids[rows.cpu().numpy()])
    chk("c1: torch.equal(adv, adv_all[rows])", torch.equal(adv, adv_all[rows]))
    part = skyrl_b200.policy_update_step(hidden[rows], weight, labels[rows], olp[rows], adv, cfg, loss_mask=mask[rows],
                                         sync=False)
    stats = sdist.all_reduce_packed_stats(part.stats)          # mean of per-rank means
    sdist.all_reduce_d_weight(part.d_weight)                   # mean of per-rank gradients (in place)
    from skyrl_b200 import _lib

    chk("c2: abs(stats[_lib.STAT_FINAL_LOSS].item() - whole.status['final", abs(stats[_lib.STAT_FINAL_LOSS].item() - whole.status["final_loss"]) < 1e-5 * max(1.0, abs(whole.status["final_loss"])))
    rel = lambda x, y: float((x.float() - y.float()).norm() / y.float().norm())
    chk("c3: rel(part.d_weight, whole.d_weight) < 2e-3", rel(part.d_weight, whole.d_weight) < 2e-3)      # bf16 d_logits are rounded after different 1/n scales
    chk("c4: rel(part.d_hidden.float() / world, whole.d_hidden[rows].floa", rel(part.d_hidden.float() / world, whole.d_hidden.view(B, A, H)[rows].reshape(-1, H).float()) < 2e-3)
    chk("c5: torch.equal(part.log_probs, whole.log_probs[rows])", torch.equal(part.log_probs, whole.log_probs[rows]))

    # ---- ragged masks: the shards hold DIFFERENT token counts, so the mean of per-rank token means is NOT the global
    # token mean.  The reference's data-parallel semantics are the former (every rank normalises by its local counts,
    # FSDP averages the gradients: worker.py:909-914, SURVEY trap 11) - assert exactly that, against the oracle.
    lens = torch.randint(A // 6, A + 1, (B,), generator=g)
    lens[: B // 4] = A                                          # a few full-length rows: the shards stay unbalanced
    rmask = (torch.arange(A)[None] < lens[:, None]).float().to(dev)
    shards = sdist.shard_prompt_groups(uids, [A] * B, world)   # whole prompt groups; same partition on every rank
    counts = [float(rmask[torch.tensor(r, device=dev)].sum()) for r in shards]
    chk("c6: len(set(counts)) > 1", len(set(counts)) > 1)                            # the test is only meaningful if the counts differ
    adv_r, _ = skyrl_b200.compute_grpo_outcome_advantage(rewards, rmask, uids)
    part = skyrl_b200.policy_update_step(hidden[rows], weight, labels[rows], olp[rows], adv_r[rows], cfg,
                                         loss_mask=rmask[rows], sync=False)
    stats = sdist.all_reduce_packed_stats(part.stats)
    # expected: every shard's own oracle loss (CPU restatement of the reference) on the step's log-probs, averaged
    local_losses, local_dw = [], []
    for r in shards:
        rr = torch.tensor(r, device=dev)
        loc = skyrl_b200.policy_update_step(hidden[rr], weight, labels[rr], olp[rr], adv_r[rr], cfg,
                                            loss_mask=rmask[rr])
        want, _ = O.loss_assembly(loc.log_probs.cpu(), olp[rr].cpu(), adv_r[rr].cpu(), cfg, rmask[rr].cpu(), None,
                                  loc.entropy.cpu())
        chk("c7: abs(loc.status['final_loss'] - want.item()) < 1e-5 * max(1.0", abs(loc.status["final_loss"] - want.item()) < 1e-5 * max(1.0, abs(want.item())))
        local_losses.append(want.item())
        local_dw.append(loc.d_weight.clone())
    mean_of_means = sum(local_losses) / world
    whole_r = skyrl_b200.policy_update_step(hidden, weight, labels, olp, adv_r, cfg, loss_mask=rmask)
    got = stats[_lib.STAT_FINAL_LOSS].item()
    chk("c8: abs(got - mean_of_means) < 1e-5 * max(1.0, abs(mean_of_means", abs(got - mean_of_means) < 1e-5 * max(1.0, abs(mean_of_means)))
    chk("c9: abs(whole_r.status['final_loss'] - mean_of_means) > 1e-4 * a", abs(whole_r.status["final_loss"] - mean_of_means) > 1e-4 * abs(mean_of_means))  # ... and NOT the global mean
    sdist.all_reduce_d_weight(part.d_weight)
    chk("c10: rel(part.d_weight, sum(local_dw) / world) < 1e-5", rel(part.d_weight, sum(local_dw) / world) < 1e-5)
    shard_w, _ = sdist.reduce_scatter_d_weight(sum(local_dw).clone())          # FSDP: mean reduce-scatter
    per = V // world
    chk("c11: rel(shard_w, sum(local_dw)[rank * per:(rank + 1) * per]) < 1", rel(shard_w, sum(local_dw)[rank * per:(rank + 1) * per]) < 1e-6)   # every rank fed the same sum

    # ---- GAE: rows are local, the whitening statistics are global (3 moments all-reduced)
    a_loc, r_loc = sdist.gae_advantage_sharded(rewards[rows], values[rows], mask[rows], 1.0, 0.95)
    chk("c12: torch.equal(r_loc, ret_all[rows]) and (a_loc - gae_all[rows]", torch.equal(r_loc, ret_all[rows]) and (a_loc - gae_all[rows]).abs().max().item() < 1e-5)

    flag = torch.tensor([0.0 if failed else 1.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DP_OK" if flag.item() == 1.0 else "DP_FAIL")
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
