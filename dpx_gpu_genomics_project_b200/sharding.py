"""Multi-GPU mode A: batches of independent pairs, length-bucketed and dealt across ranks.

One process per GPU (torch.distributed; NCCL on the GPU box, gloo in the CPU tests).  The alignment of a
pair never needs another pair, so there is NO data-path collective: every rank aligns its own shard with its
own device context; only the small result arrays (score, end cell) travel back to rank 0, where they are put
back in pair order.  The reference has nothing comparable (single device 0 only,
cuda/LNW/LinearNeedlemanWunschV19.cu:362-376 merely prints the device count).
"""
from __future__ import annotations

import numpy as np


def shard_indices(pairs: np.ndarray, world_size: int) -> list[np.ndarray]:
    """Deal pairs to ranks so that every rank gets (almost) the same number of DP cells.

    Pairs are sorted by cell count Q*R (descending, ties by index: deterministic) and dealt in boustrophedon
    order (0..W-1, W-1..0, ...), the classic longest-processing-time heuristic for equal bins; within a rank the
    indices are returned ascending so that the blob ranges a rank touches stay mostly contiguous."""
    n = len(pairs)
    if world_size <= 1:
        return [np.arange(n, dtype=np.int64)]
    cells = pairs["referenceSize"].astype(np.int64) * pairs["querySize"].astype(np.int64)
    order = np.lexsort((np.arange(n), -cells))
    k = np.arange(n)
    rnd, pos = k // world_size, k % world_size
    owner_sorted = np.where(rnd % 2 == 0, pos, world_size - 1 - pos)
    owner = np.empty(n, dtype=np.int64)
    owner[order] = owner_sorted
    return [np.flatnonzero(owner == r) for r in range(world_size)]


def shard_cells(pairs: np.ndarray, shards: list[np.ndarray]) -> np.ndarray:
    cells = pairs["referenceSize"].astype(np.int64) * pairs["querySize"].astype(np.int64)
    return np.array([int(cells[s].sum()) for s in shards], dtype=np.int64)


def align_sharded(compute, sequences: np.ndarray, pairs: np.ndarray, rank: int, world_size: int, dist=None):
    """Aligns `pairs` across `world_size` ranks.  `compute(sequences, pairs_subset) -> (scores int32[m], end_rc int32[m,2])`
    runs on this rank's device (Engine.align_batch in production).  Returns (scores, end_rc) in ORIGINAL pair order on
    rank 0 and (None, None) elsewhere.  The only communication is the gather of the result arrays."""
    shards = shard_indices(pairs, world_size)
    mine = shards[rank]
    s, e = compute(sequences, pairs[mine])
    s = np.ascontiguousarray(s, dtype=np.int32)
    e = np.ascontiguousarray(e, dtype=np.int32).reshape(-1, 2)
    if world_size == 1 or dist is None:
        scores = np.empty(len(pairs), np.int32); end_rc = np.empty((len(pairs), 2), np.int32)
        scores[mine] = s; end_rc[mine] = e
        return scores, end_rc
    import torch
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    m = max(len(x) for x in shards)
    buf = torch.zeros((m, 3), dtype=torch.int32, device=dev)
    if len(mine):
        buf[: len(mine), 0] = torch.from_numpy(s).to(dev)
        buf[: len(mine), 1:] = torch.from_numpy(e).to(dev)
    gathered = [torch.zeros_like(buf) for _ in range(world_size)] if rank == 0 else None
    if backend == "nccl":
        # NCCL has no gather on every torch build: all_gather of the (small) result table is equivalent here
        tmp = [torch.zeros_like(buf) for _ in range(world_size)]
        dist.all_gather(tmp, buf)
        gathered = tmp
    else:
        dist.gather(buf, gathered, dst=0)
    if rank != 0:
        return None, None
    scores = np.empty(len(pairs), np.int32); end_rc = np.empty((len(pairs), 2), np.int32)
    for r in range(world_size):
        g = gathered[r].cpu().numpy()
        scores[shards[r]] = g[: len(shards[r]), 0]
        end_rc[shards[r]] = g[: len(shards[r]), 1:]
    return scores, end_rc
