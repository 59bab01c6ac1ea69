"""Multi-GPU plumbing of the policy update: one process per GPU, data parallel over whole prompt groups.

What shards: rollout batches split by *whole prompt groups* (all samples of a uid on one rank), so
GRPO / RLOO statistics stay local and the per-token loss needs no exchange.  What crosses NVLink:

1. ``all_reduce_metrics`` — the status dict of a micro-batch, all keys in ONE collective (the reference issues
   one NCCL all-reduce + H2D + D2H per scalar: workers/worker_utils.py:25-35, distributed/strategy.py:70-95).
   Same semantics: mean across ranks, ``*_min`` -> min, ``*_max`` -> max.
2. ``all_reduce_d_weight`` — lm_head weight-gradient all-reduce (mean, fp32: FSDP ``reduce_dtype=fp32``,
   fsdp_strategy.py:218-226), once per optimizer step, on a side stream.
3. ``gae_advantage_sharded`` — GAE whitening needs global moments (trainer.py:268-276): 3 doubles all-reduced
   between the scan and the whiten kernels.

Works with the ``nccl`` backend on GPUs and with ``gloo`` on CPU tensors (used by the CPU tests for the host logic).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_prompt_groups(index: Sequence, valid_tokens: Sequence[int], world_size: int) -> List[List[int]]:
    """Assign whole uid groups to ranks, balancing valid-token counts (longest-processing-time first).

    Returns ``rows_per_rank[r]`` = sorted row indices for rank r.  Deterministic: ties break on first
    appearance of the uid, so every rank computes the same partition from the same inputs.
    """
    groups: Dict = {}
    order: List = []
    for i, u in enumerate(index):
        u = u.item() if isinstance(u, np.generic) else u
        if u not in groups:
            groups[u] = []
            order.append(u)
        groups[u].append(i)
    weights = [(sum(int(valid_tokens[i]) for i in groups[u]), k, u) for k, u in enumerate(order)]
    weights.sort(key=lambda x: (-x[0], x[1]))
    load = [0] * world_size
    rows: List[List[int]] = [[] for _ in range(world_size)]
    for wgt, _, u in weights:
        r = min(range(world_size), key=lambda j: (load[j], j))
        load[r] += wgt
        rows[r].extend(groups[u])
    return [sorted(r) for r in rows]


def reduce_metrics(metrics: Dict[str, List[float]]) -> Dict[str, float]:
    """Fold per-micro-batch metric lists: mean, ``*_max`` -> max, ``*_min`` -> min (worker_utils.py:8-22)."""
    out = {}
    for k, v in metrics.items():
        assert len(v) > 0, f"No metrics for key {k}"
        assert all(isinstance(x, (int, float)) for x in v), f"Metrics for key {k} are not all numbers"
        if k.endswith("_max"):
            out[k] = max(v)
        elif k.endswith("_min"):
            out[k] = min(v)
        else:
            out[k] = sum(v) / len(v)
    return out


def _world(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def all_reduce_metrics(status: Dict[str, float], device: Optional[torch.device] = None, group=None) -> Dict[str, float]:
    """Reference semantics of ``all_reduce_metrics`` (worker_utils.py:25-35) in a single collective.

    Every rank contributes one packed fp32 vector; one ``all_gather`` moves it; mean / min / max are folded
    locally in rank order (deterministic).  Non-numeric entries are passed through untouched.
    """
    keys = [k for k, v in status.items() if isinstance(v, (int, float))]
    ws = _world(group)
    if ws == 1 or not keys:
        return dict(status)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    vec = torch.tensor([float(status[k]) for k in keys], dtype=torch.float32, device=device)
    gathered = torch.empty(ws * len(keys), dtype=torch.float32, device=device)
    dist.all_gather_into_tensor(gathered, vec, group=group)
    table = gathered.view(ws, len(keys)).cpu()
    out = dict(status)
    for j, k in enumerate(keys):
        col = table[:, j]
        if k.endswith("_max"):
            out[k] = float(col.max())
        elif k.endswith("_min"):
            out[k] = float(col.min())
        else:
            # the reference divides each rank's value by world_size and SUM-reduces
            out[k] = float((col / ws).sum())
    return out


def all_reduce_packed_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """Mean-of-local-means of a packed device stats vector (one NCCL all-reduce, no host round trip)."""
    ws = _world(group)
    if ws == 1:
        return stats
    out = stats / ws
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


def all_reduce_d_weight(d_weight: torch.Tensor, group=None, async_op: bool = False):
    """Mean all-reduce of the fp32 lm_head weight gradient over NVLink (FSDP averages gradients over DP ranks).

    On NCCL the averaging happens inside the collective (``ReduceOp.AVG``): no separate 1/world pass over the
    2.18 GB gradient.  (gloo has no AVG: scale, then SUM.)"""
    ws = _world(group)
    if ws == 1:
        return None
    if d_weight.is_cuda and dist.get_backend(group) == "nccl":
        return dist.all_reduce(d_weight, op=dist.ReduceOp.AVG, group=group, async_op=async_op)
    d_weight.div_(ws)
    return dist.all_reduce(d_weight, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def reduce_scatter_d_weight(d_weight: torch.Tensor, out: Optional[torch.Tensor] = None, group=None,
                            async_op: bool = False):
    """What FSDP actually does with the gradient of a sharded parameter (fsdp_strategy.py:151-153, 214-226: fp32
    ``reduce_dtype``, gradients averaged over the data-parallel ranks): a mean REDUCE-SCATTER - rank r keeps rows
    ``[r * V / world, (r + 1) * V / world)`` of the averaged ``[V, H]`` gradient, half the bytes of an all-reduce.

    Returns ``(shard, work)``; ``shard`` is ``out`` or a new ``[V / world, H]`` tensor.  ``V`` must divide evenly
    (Qwen vocabularies do for 2 / 4 / 8 ranks; pad the head otherwise)."""
    ws = _world(group)
    if ws == 1:
        return d_weight, None
    V = d_weight.shape[0]
    if V % ws != 0:
        raise ValueError(f"reduce_scatter_d_weight needs V ({V}) divisible by the group size ({ws})")
    if out is None:
        out = torch.empty((V // ws,) + tuple(d_weight.shape[1:]), dtype=d_weight.dtype, device=d_weight.device)
    if d_weight.is_cuda and dist.get_backend(group) == "nccl":
        work = dist.reduce_scatter_tensor(out, d_weight, op=dist.ReduceOp.AVG, group=group, async_op=async_op)
        return out, work
    # gloo (CPU tests of the host logic) has neither AVG nor reduce-scatter: all-reduce a scaled copy, keep the shard
    full = d_weight / ws
    dist.all_reduce(full, op=dist.ReduceOp.SUM, group=group)
    r = dist.get_rank(group)
    out.copy_(full[r * (V // ws):(r + 1) * (V // ws)])
    return out, None


class DWeightScatter:
    """Peer-mapped gradient shards for the fused ``d_weight`` GEMM + reduce-scatter (``policy_update_step(...,
    dw_scatter=...)``): what FSDP's post-backward reduce-scatter gives the reference (fsdp_strategy.py:151-153,
    214-226 - fp32, averaged over the data-parallel ranks, rank r keeps rows ``[r V / world, (r + 1) V / world)``),
    produced by the GEMM epilogue itself over NVLink instead of by a collective after the step.

    Construction is collective over ``group`` (one process per GPU on one node): every rank allocates its
    ``[V / world, H]`` fp32 shard, exports a CUDA IPC handle and opens the other ranks' shards.

    Per optimizer step::

        sc.begin()                      # zero my shard; barrier: nobody adds into a shard before it is zeroed
        policy_update_step(..., dw_scatter=sc)   # any number of micro-batches: contributions accumulate
        sc.finish()                     # barrier: every rank's kernels are done -> sc.shard is complete

    Both barriers are one-element NCCL all-reduces in stream order (no host sync).  The sum order of the ranks'
    contributions is not fixed, so results are reproducible to fp32 rounding only; ``reduce_scatter_d_weight`` on a
    local gradient is the bit-reproducible alternative.
    """

    def __init__(self, V: int, H: int, group=None, device: Optional[torch.device] = None, remote: str = "tma"):
        """``remote``: how the peers' pieces travel - ``"tma"`` (TMA reduce-add on a tensor map over the peer-mapped
        shard) or ``"red"`` (16-byte fp32 vector reductions issued by the epilogue threads; the fallback)."""
        import ctypes as C

        from . import _lib

        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("DWeightScatter needs an initialised torch.distributed process group")
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if not 2 <= self.world <= 8:
            raise ValueError(f"DWeightScatter supports 2..8 ranks of one node, got {self.world}")
        if V % (32 * self.world) != 0:
            raise ValueError(f"V ({V}) must be a multiple of 32 * world ({32 * self.world})")
        self.shape = (V, H)
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        lib = _lib.load()
        with torch.cuda.device(self.device):
            self.shard = torch.zeros((V // self.world, H), dtype=torch.float32, device=self.device)
            self._flag = torch.zeros(1, dtype=torch.float32, device=self.device)
            handle = (C.c_ubyte * 64)()
            off = C.c_int64(0)
            _lib.check(lib.skyb200_ipc_export(_lib.ptr(self.shard), handle, C.byref(off)), "skyb200_ipc_export")
            mine = (bytes(handle), int(off.value))
            everyone = [None] * self.world
            dist.all_gather_object(everyone, mine, group=group)
            self._opened = []
            self.c_struct = _lib.DwScatter()
            self.c_struct.world, self.c_struct.rank = self.world, self.rank
            self.c_struct.scale = 1.0 / self.world
            if remote not in ("tma", "red"):
                raise ValueError("remote must be 'tma' or 'red'")
            self.c_struct.remote_red = 1 if remote == "red" else 0
            for r, (h, o) in enumerate(everyone):
                if r == self.rank:
                    self.c_struct.shard[r] = self.shard.data_ptr()
                    continue
                p = C.c_void_p()
                buf = (C.c_ubyte * 64).from_buffer_copy(h)
                _lib.check(lib.skyb200_ipc_open(buf, o, C.byref(p)), "skyb200_ipc_open")
                self._opened.append(p.value)
                self.c_struct.shard[r] = p.value
        self._barrier()

    def _barrier(self):
        dist.all_reduce(self._flag, op=dist.ReduceOp.SUM, group=self.group)

    def begin(self):
        """Zero this rank's shard, then a stream-ordered barrier (no peer adds into it before the zero)."""
        self.shard.zero_()
        self._barrier()

    def finish(self) -> torch.Tensor:
        """Stream-ordered barrier after the last step: all ranks' kernels (and their NVLink reductions) are done."""
        self._barrier()
        return self.shard

    def close(self):
        from . import _lib

        if getattr(self, "_opened", None):
            torch.cuda.synchronize(self.device)
            self._barrier()  # nobody is still adding into a shard that is about to be unmapped / freed
            torch.cuda.synchronize(self.device)
            lib = _lib.load()
            with torch.cuda.device(self.device):
                for p in self._opened:
                    lib.skyb200_ipc_close(p)
            self._opened = []


def gae_advantage_sharded(token_level_rewards, values, response_mask, gamma: float, lambd: float, group=None):
    """GAE on this rank's rows with GLOBAL whitening statistics (3 doubles all-reduced).

    Equals ``compute_gae_advantage_return`` on the concatenated batch (trainer.py:268-276 computes advantages
    over the whole batch before sharding).
    """
    from . import _lib
    from .advantages import _f32, _mask_arg

    lib = _lib.load()
    dev = token_level_rewards.device
    if not token_level_rewards.is_cuda:
        raise RuntimeError("gae_advantage_sharded needs CUDA tensors")
    B, A = token_level_rewards.shape
    with torch.cuda.device(dev):
        r, v = _f32(token_level_rewards, dev), _f32(values, dev)
        m, mtype = _mask_arg(response_mask, dev)
        adv = torch.empty((B, A), dtype=torch.float32, device=dev)
        ret = torch.empty((B, A), dtype=torch.float32, device=dev)
        moments = torch.zeros(3, dtype=torch.float64, device=dev)
        stats = torch.empty(4, dtype=torch.float32, device=dev)
        wsb = lib.skyb200_gae_workspace(B)
        ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
        _lib.check(lib.skyb200_gae_scan(_lib.ptr(r), _lib.ptr(v), _lib.ptr(m), mtype, _lib.ptr(adv), _lib.ptr(ret),
                                        _lib.ptr(moments), _lib.ptr(ws), wsb, B, A, float(gamma), float(lambd),
                                        _lib.stream_ptr()), "skyb200_gae_scan")
        if _world(group) > 1:
            dist.all_reduce(moments, op=dist.ReduceOp.SUM, group=group)
        _lib.check(lib.skyb200_whiten_from_moments(_lib.ptr(adv), _lib.ptr(moments), _lib.ptr(stats), B * A,
                                                   _lib.stream_ptr()), "skyb200_whiten_from_moments")
        n = float(stats[0].item())
    if n == 0:
        raise ValueError("At least one element in the mask has to be 1.")
    if n == 1:
        raise ValueError("The sum of the mask is one, which can cause a division by zero.")
    return adv, ret
