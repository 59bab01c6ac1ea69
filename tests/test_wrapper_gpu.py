"""The wrapper call-site glue (skyrl_b200/model_wrapper.py) against the reference wrapper's arithmetic
(skyrl-train/skyrl_train/model_wrapper.py:261-375) restated with the CPU oracle: project EVERY position, roll the
labels, (pad back,) slice ``[:, -A-1:-1]``."""
import pytest
import torch

import skyrl_b200
from oracle import policy_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _inputs(B, S, H, V, seed, left_pad=True):
    g = torch.Generator().manual_seed(seed)
    sequences = torch.randint(1, V, (B, S), generator=g)
    am = torch.ones(B, S, dtype=torch.long)
    for b in range(B):
        lp = int(torch.randint(0, S // 4, (1,), generator=g)) if left_pad else 0
        rp = int(torch.randint(0, S // 3, (1,), generator=g))
        am[b, :lp] = 0
        if rp:
            am[b, S - rp:] = 0
    am[0] = 1  # one sequence without any padding
    weight = (torch.randn(V, H, generator=g) * (1.5 / H ** 0.5)).bfloat16()
    return g, sequences, am, weight


def _reference_padded(hidden, weight, sequences, am, A, T, with_entropy):
    """model_wrapper.py:291, 309-321, 338-349, 370 without sample packing."""
    B, S, H = hidden.shape
    lp, ent = O.head_logprobs_entropy(hidden.reshape(-1, H), weight, torch.roll(sequences, -1, 1).reshape(-1), T)
    lp, ent = lp.view(B, S), ent.view(B, S) * am
    return lp[:, -A - 1:-1], (ent[:, -A - 1:-1] if with_entropy else None)


@pytest.mark.parametrize("A", [1, 37, 63])
def test_padded_wrapper_matches_reference_slicing(A):
    B, S, H, V, T = 4, 64, 64, 1000, 0.8
    g, sequences, am, weight = _inputs(B, S, H, V, 3)
    hidden = torch.randn(B, S, H, generator=g).bfloat16()
    h_ref = hidden.float().requires_grad_(True)
    lp_ref, ent_ref = _reference_padded(h_ref, weight.float(), sequences, am, A, T, True)
    gl = torch.randn(B, A, generator=g)
    (lp_ref * gl).sum().backward()
    h = hidden.to(DEV).requires_grad_(True)
    lp, ent = skyrl_b200.action_log_probs_from_hidden(h, weight.to(DEV), sequences.to(DEV), A,
                                                      attention_mask=am.to(DEV), temperature=T, compute_entropy=True)
    assert lp.shape == (B, A) and ent.shape == (B, A)
    assert (lp.cpu() - lp_ref.detach()).abs().max().item() < 1e-3
    assert (ent.cpu() - ent_ref.detach()).abs().max().item() < 1e-3
    (lp * gl.to(DEV)).sum().backward()
    # rows outside the window never reach the head: their gradient is exactly zero, like the reference's slice
    assert torch.all(h.grad[:, :S - A - 1] == 0) and torch.all(h.grad[:, S - 1:] == 0)
    rel = (h.grad.float().cpu() - h_ref.grad).norm() / h_ref.grad.norm()
    assert rel.item() < 2e-2


@pytest.mark.parametrize("A", [5, 40])
def test_sample_packing_wrapper_matches_unpad_project_pad(A):
    """model_wrapper.py:278-289 (unpad), :291 (roll of the PACKED stream), :353-358 (pad_input -> zeros), :370."""
    B, S, H, V, T = 5, 64, 64, 1000, 1.0
    g, sequences, am, weight = _inputs(B, S, H, V, 9)
    valid = am.reshape(-1) != 0
    idx = torch.nonzero(valid).squeeze(-1)
    nnz = idx.numel()
    packed_hidden = torch.randn(nnz, H, generator=g).bfloat16()   # what the trunk returns for the packed stream
    packed_tokens = sequences.reshape(-1)[idx]
    h_ref = packed_hidden.float().requires_grad_(True)
    lp_p, ent_p = O.head_logprobs_entropy(h_ref, weight.float(), torch.roll(packed_tokens, -1, 0), T)
    pad = lambda x: torch.zeros(B * S).index_copy(0, idx, x).view(B, S)   # pad_input
    lp_ref, ent_ref = pad(lp_p)[:, -A - 1:-1], pad(ent_p)[:, -A - 1:-1]
    gl = torch.randn(B, A, generator=g)
    (lp_ref * gl).sum().backward()
    h = packed_hidden.to(DEV).requires_grad_(True)
    lp, ent = skyrl_b200.action_log_probs_from_packed_hidden(h.unsqueeze(0), weight.to(DEV), sequences.to(DEV),
                                                             am.to(DEV), A, temperature=T, compute_entropy=True)
    assert lp.shape == (B, A)
    assert (lp.cpu() - lp_ref.detach()).abs().max().item() < 1e-3
    assert (ent.cpu() - ent_ref.detach()).abs().max().item() < 1e-3
    assert torch.equal(lp.cpu() == 0, lp_ref.detach() == 0)      # the padding of the window stays exactly zero
    (lp * gl.to(DEV)).sum().backward()
    rel = (h.grad.float().cpu() - h_ref.grad).norm() / h_ref.grad.norm()
    assert rel.item() < 2e-2
    with pytest.raises(ValueError):
        skyrl_b200.action_log_probs_from_packed_hidden(h[:-1], weight.to(DEV), sequences.to(DEV), am.to(DEV), A)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (NCCL sequence-parallel group)")
def test_sequence_parallel_glue_over_nccl_two_ranks():
    """SP = 2 equals SP = 1 with the real fused head on two GPUs (the reference's own check:
    tests/gpu/gpu_ci/test_model_wrapper.py:134-191); windows inside one slice, across both slices, and A = S - 1."""
    import os
    import subprocess
    import sys

    script = os.path.join(os.path.dirname(__file__), "sp_nccl_worker.py")
    env = dict(os.environ, PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29547", script], env=env, capture_output=True,
                       text=True, timeout=280)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SP_OK" in r.stdout
