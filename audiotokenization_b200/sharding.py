"""Utterance sharding across ranks (one process per GPU) and host-side gathering.

The path has no cross-utterance state (no batch-norm; LSTM state is per item), so the
utterance list is partitioned statically and each rank runs independently
(SURVEY.md section 8e).  There is NO data-path collective: the int16 index arrays
(160 B per audio-second) are gathered on the host through a CPU (gloo) group.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def shard_range(num_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced [start, end) slice of ``num_items`` for ``rank`` (sizes differ by at most 1)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, extra = divmod(num_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_by_cost(costs: Sequence[int], world_size: int) -> List[List[int]]:
    """Greedy longest-processing-time partition for ragged utterance lengths: returns, per rank,
    the (ascending) item ids it owns.  Deterministic (ties broken by item id / rank id)."""
    order = sorted(range(len(costs)), key=lambda i: (-int(costs[i]), i))
    loads = [0] * world_size
    owned: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        owned[r].append(i)
        loads[r] += int(costs[i])
    return [sorted(o) for o in owned]


def gather_indices_to_rank0(local: np.ndarray, group: Optional["dist.ProcessGroup"] = None) -> Optional[np.ndarray]:
    """Concatenate each rank's ``[n_local, T', n_q]`` int16 block on rank 0 in rank order (host memory,
    CPU process group).  Returns None on other ranks.  Without an initialised process group it is the identity."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    local = np.ascontiguousarray(local)
    if local.ndim > 4:
        raise ValueError("gather_indices_to_rank0: at most 4 dimensions")
    # 1. everyone learns every shard's shape (gloo gathers need equal sizes, so pad the payload to the max)
    shape = torch.tensor(list(local.shape) + [-1] * (4 - local.ndim), dtype=torch.int64)
    shapes = [torch.zeros(4, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(shapes, shape, group=group)
    nbytes = [int(np.prod([int(v) for v in s_ if int(v) >= 0])) * local.itemsize for s_ in shapes]
    payload = torch.zeros(max(max(nbytes), 1), dtype=torch.uint8)
    payload[: local.nbytes] = torch.from_numpy(local.reshape(-1).view(np.uint8))
    bufs = [torch.empty_like(payload) for _ in range(world)] if rank == 0 else None
    dist.gather(payload, bufs, dst=0, group=group)
    if rank != 0:
        return None
    parts = []
    for b, s_, nb in zip(bufs, shapes, nbytes):
        shp = [int(v) for v in s_ if int(v) >= 0]
        parts.append(b[:nb].numpy().view(local.dtype).reshape(shp))
    return np.concatenate(parts, axis=0)
