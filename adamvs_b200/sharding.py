"""Multi-GPU plumbing for the one place the path shards: whole reference views.

A scene is a list of reference views; each view (with its source views) is an independent unit, so
ranks take contiguous slices of the list and never exchange data on the hot path.  The only
communication is host-side: a barrier around timed regions, a MAX over ranks of device times, and
the final gather of per-view results to rank 0 — all through ``torch.distributed`` (NCCL on the GPU
box, gloo in the CPU tests).  The reference's own mechanism is nn.DataParallel's batch scatter
(predict_whu.py:82), which degenerates to one GPU at its default batch size of 1.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_units: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of ``range(n_units)``: the first ``n_units % world`` ranks get
    one extra unit.  Returns [begin, end)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(int(n_units), world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def batches(begin: int, end: int, batch: int) -> List[Tuple[int, int]]:
    """Split one rank's slice into launches of at most ``batch`` views (the last may be ragged)."""
    return [(i, min(i + batch, end)) for i in range(begin, end, max(1, batch))]


def world_info() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def max_over_ranks(value: float, device: Optional[torch.device] = None) -> float:
    """MAX all-reduce of a host scalar (device times are compared this way, never wall clocks)."""
    _, world = world_info()
    if world == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_views(local: Sequence[torch.Tensor], n_units: int, dst: int = 0) -> Optional[List[torch.Tensor]]:
    """Gather per-view CPU tensors (e.g. depth maps [H,W]) from every rank's contiguous slice to
    ``dst`` in global view order.  Returns the full list on ``dst`` and None elsewhere."""
    rank, world = world_info()
    begin, end = shard_range(n_units, rank, world)
    if len(local) != end - begin:
        raise ValueError(f"rank {rank} holds {len(local)} views, its slice [{begin},{end}) needs {end - begin}")
    local = [t.detach().cpu() for t in local]
    if world == 1:
        return list(local)
    gathered = [None] * world if rank == dst else None
    dist.gather_object(local, gathered, dst=dst)
    if rank != dst:
        return None
    out: List[torch.Tensor] = []
    for r, part in enumerate(gathered):
        b, e = shard_range(n_units, r, world)
        assert len(part) == e - b
        out.extend(part)
    return out


def run_sharded(n_units: int, batch: int, run_batch: Callable[[int, int], Sequence[torch.Tensor]], dst: int = 0):
    """Drive ``run_batch(begin, end) -> per-view tensors`` over this rank's slice and gather on ``dst``."""
    rank, world = world_info()
    begin, end = shard_range(n_units, rank, world)
    local: List[torch.Tensor] = []
    for b, e in batches(begin, end, batch):
        res = run_batch(b, e)
        if len(res) != e - b:
            raise ValueError("run_batch must return one tensor per view")
        local.extend(res)
    return gather_views(local, n_units, dst)
