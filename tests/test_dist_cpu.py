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
        # FSDP semantics: mean reduce-scatter, rank r keeps rows [r * V / world, (r + 1) * V / world)
        full = torch.arange(12, dtype=torch.float32).view(4, 3) * (rank + 1)
        shard, _ = sdist.reduce_scatter_d_weight(full)
        want = (torch.arange(12, dtype=torch.float32).view(4, 3) * 1.5)[2 * rank:2 * rank + 2]
        assert torch.allclose(shard, want), (shard, want)
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


@pytest.mark.timeout(180)
@pytest.mark.parametrize("world", [2, 3])
def test_sequence_parallel_wrapper_glue(world, tmp_path):
    """action_log_probs_from_sp_hidden (Ulysses branch, model_wrapper.py:292-355) over gloo with a stand-in head:
    every rank ends up with the single-process [B, A] log-probs and entropies, the gradient of its own hidden slice
    (times the group size with the reference's grad_scaler, :229-231), and the weight gradients add up - including
    sequence lengths that need padding and ranks whose slice holds no response position at all."""
    import subprocess
    import sys

    port = str(_free_port())
    out = str(tmp_path / "sp")
    worker = os.path.join(os.path.dirname(__file__), "sp_worker.py")
    procs = [subprocess.Popen([sys.executable, worker, str(r), str(world), port, out]) for r in range(world)]
    for p in procs:
        assert p.wait(timeout=150) == 0
    empty_slices = 0
    for r in range(world):
        res = torch.load(f"{out}.{r}")
        for case, v in res.items():
            assert v["fwd"] and v["dh"] and v["dw"], (world, r, case, v)
            empty_slices += v["rows_here"] == 0
    assert empty_slices > 0  # the zero-anchor path (a rank without response rows still back-propagates) was taken
