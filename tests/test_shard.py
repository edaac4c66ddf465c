"""Host-side multi-GPU logic on CPU: LPT partition + result all-gather over gloo, world_size 2."""
import os
import socket
import zlib

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import product


def test_partition_is_balanced_and_deterministic():
    shard = product().shard
    sizes = [100, 1, 1, 1, 50, 49, 2, 98]
    p = shard.partition(sizes, 2)
    assert sorted(p[0] + p[1]) == list(range(8))
    loads = [sum(sizes[i] for i in s) for s in p]
    assert abs(loads[0] - loads[1]) <= 2
    assert p == shard.partition(sizes, 2)
    assert shard.partition([], 4) == [[], [], [], []]
    assert shard.partition([5], 3)[0] == [0]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import importlib
    shard = importlib.import_module("swift-png_b200").shard
    streams = [zlib.compress(bytes([i]) * (1000 * (i + 1))) for i in range(9)]

    def work(idxs):  # stand-in for the per-GPU decode: inflate on the host
        out = []
        for i in idxs:
            data = zlib.decompress(streams[i])
            out.append((0, zlib.adler32(data), len(data)))
        return out

    res = shard.run_sharded([len(s) for s in streams], work)
    # the benchmark's layout: rank-major, the same number of jobs on every rank, no LPT
    eq_sizes = [7, 8, 9, 7, 8, 9]
    eq = shard.run_sharded(eq_sizes, lambda idxs: [(0, 100 * rank + i, eq_sizes[i]) for i in idxs], equal_shards=3)
    q.put((rank, (res, eq)))
    dist.destroy_process_group()


def test_sharded_run_over_gloo_world2():
    import importlib
    importlib.import_module("swift-png_b200")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = [(0, zlib.adler32(bytes([i]) * (1000 * (i + 1))), 1000 * (i + 1)) for i in range(9)]
    assert got[0][0] == expect and got[1][0] == expect
    eq = [(0, 0, 7), (0, 1, 8), (0, 2, 9), (0, 103, 7), (0, 104, 8), (0, 105, 9)]
    assert got[0][1] == eq and got[1][1] == eq
