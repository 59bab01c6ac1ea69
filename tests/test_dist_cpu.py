"""World-size-2 gloo tests of the multi-GPU host logic (packed metric reduction, group sharding equivalence)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from skyrl_b200 import dist as sdist

        status = {"final_loss": 1.0 + rank, "policy_loss": 2.0 * (rank + 1), "loss_metrics/is_ratio_max": 3.0 + rank,
                  "loss_metrics/is_ratio_min": 0.5 - rank, "response_length": 7, "note": "keep"}
        red = sdist.all_reduce_metrics(status, device=torch.device("cpu"))
        stats = torch.tensor([1.0 + rank, 10.0 * (rank + 1)])
        packed = sdist.all_reduce_packed_stats(stats)
        dw = torch.full((4, 3), float(rank + 1))
        sdist.all_reduce_d_weight(dw)
        q.put((rank, red, packed.tolist(), dw[0, 0].item()))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_all_reduce_metrics_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=100) for _ in range(2)]
    for p in procs:
        p.join(30)
        assert p.exitcode == 0
    for rank, red, packed, dw in results:
        # reference semantics (worker_utils.py:25-35): mean, *_max -> max, *_min -> min
        assert red["final_loss"] == pytest.approx(1.5)
        assert red["policy_loss"] == pytest.approx(3.0)
        assert red["loss_metrics/is_ratio_max"] == pytest.approx(4.0)
        assert red["loss_metrics/is_ratio_min"] == pytest.approx(-0.5)
        assert red["response_length"] == pytest.approx(7.0)
        assert red["note"] == "keep"
        assert packed == pytest.approx([1.5, 15.0])
        assert dw == pytest.approx(1.5)  # mean of the per-rank gradients


def test_single_process_is_identity():
    from skyrl_b200 import dist as sdist

    s = {"a": 1.0, "b_max": 2.0}
    assert sdist.all_reduce_metrics(s) == s


def test_combine_shard_stats_is_the_log_sum_exp_merge():
    """Host logic of the vocab-parallel head on CPU tensors: combining per-shard (max, sum, weighted sum, label logit)
    equals the direct computation on the full vocabulary."""
    import math

    from skyrl_b200.tp_head import combine_shard_stats

    g = torch.Generator().manual_seed(0)
    N, V, T = 50, 300, 0.7
    logits = torch.randn(N, V, generator=g) * 3
    labels = torch.randint(0, V, (N,), generator=g)
    c = math.log2(math.e) / T
    cuts = [0, 100, 130, V]
    stats = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        y = logits[:, a:b] * c
        m = y.max(-1).values
        e = torch.exp2(y - m[:, None])
        inside = (labels >= a) & (labels < b)
        z = torch.where(inside, logits.gather(-1, labels[:, None]).squeeze(-1), torch.zeros(N))
        stats.append(torch.stack([m, e.sum(-1), (e * logits[:, a:b]).sum(-1), z], -1))
    lp, ent, _ = combine_shard_stats(torch.stack(stats), T)
    ref = torch.log_softmax(logits / T, -1)
    assert torch.allclose(lp, ref.gather(-1, labels[:, None]).squeeze(-1), atol=1e-5)
    assert torch.allclose(ent, -(ref.exp() * ref).sum(-1), atol=1e-5)
