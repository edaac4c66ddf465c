"""Multi-GPU sharding of a batch of independent images/streams.

Every image is an independent unit (the reference has no cross-image state either), so the batch
is partitioned by compressed size with longest-processing-time-first and each rank (one process
per GPU) decodes its shard with no data-path collective.  The only exchange is the all-gather of
the per-image result words (status, checksum, produced) so that every rank sees the whole batch's
outcome; decoded pixels stay on the GPU that produced them.
"""
from __future__ import annotations

import heapq
from typing import Callable, Sequence


def partition(sizes: Sequence[int], world: int) -> list[list[int]]:
    """LPT: biggest job first onto the least-loaded rank.  Deterministic (ties by index), so
    every rank computes the same answer from the same size list."""
    heap = [(0, r) for r in range(world)]
    heapq.heapify(heap)
    shards: list[list[int]] = [[] for _ in range(world)]
    for i in sorted(range(len(sizes)), key=lambda k: (-sizes[k], k)):
        load, r = heapq.heappop(heap)
        shards[r].append(i)
        heapq.heappush(heap, (load + sizes[i], r))
    for s in shards:
        s.sort()
    return shards


def run_sharded(sizes: Sequence[int], work: Callable[[list[int]], list[tuple[int, int, int]]],
                group=None, device=None, equal_shards: int = 0):
    """Each rank runs `work(my_indices)` -> [(status, checksum, produced)] and the results of the
    whole batch come back on every rank, in batch order.  `group`: a torch.distributed process
    group (NCCL on GPUs, gloo in the CPU tests); None = single process.  `equal_shards` = n: the
    batch is rank-major with n jobs per rank already (weak-scaling benchmark), no LPT needed."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if (group is not None or dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    if equal_shards:
        assert len(sizes) == equal_shards * world
        shards = [list(range(r * equal_shards, (r + 1) * equal_shards)) for r in range(world)]
    else:
        shards = partition(sizes, world)
    mine = work(shards[rank])
    assert len(mine) == len(shards[rank])
    if world == 1:
        return list(mine)
    width = max(len(s) for s in shards)
    local = torch.full((width, 3), -1, dtype=torch.int64, device=device)
    if mine:
        local[: len(mine)] = torch.tensor(mine, dtype=torch.int64, device=device)
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local, group=group)
    out: list = [None] * len(sizes)
    for r, idxs in enumerate(shards):
        rows = gathered[r].cpu().tolist()
        for k, i in enumerate(idxs):
            out[i] = tuple(rows[k])
    return out
