"""CPU multi-process tests (gloo, world_size 2 and 3) of the N>1 host logic: shard ranges and the per-step all-gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lsc_dr_planner_b200.sharding import all_agree, allgather_handles, allgather_rows, shard_range, shard_sizes


def test_shard_ranges_partition_the_agents():
    for n in (0, 1, 7, 64, 1000, 4096):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b
            assert sum(shard_sizes(n, world)) == n
            assert max(shard_sizes(n, world)) - min(shard_sizes(n, world)) <= (n + world - 1) // world


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(n_total * 5 * 6 * 3, dtype=torch.float32).view(n_total, 5, 6, 3)
        lo, hi = shard_range(n_total, rank, world)
        got = allgather_rows(full[lo:hi].clone(), n_total)
        ok = torch.equal(got, full)
        state = torch.arange(n_total * 9, dtype=torch.float32).view(n_total, 9)
        ok = ok and torch.equal(allgather_rows(state[lo:hi].clone(), n_total), state)
        # max-over-ranks timing reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ok = ok and float(t) == float(world)
        # plumbing of the peer exchange (ClosedLoopSim._connect_peers): 64-byte handles in rank order + success flags
        mine = bytes([(rank * 37 + i) % 251 for i in range(64)])
        handles, all_ok = allgather_handles(mine, ok=True)
        ok = ok and all_ok and len(handles) == 64 * world
        for r in range(world):
            ok = ok and handles[64 * r:64 * (r + 1)] == bytes([(r * 37 + i) % 251 for i in range(64)])
        _, all_ok = allgather_handles(mine, ok=(rank != world - 1))        # one rank could not create its block
        ok = ok and not all_ok and not all_agree(rank != 0) and all_agree(True)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 64), (2, 7), (3, 10)])
def test_allgather_rows_gloo(world, n_total):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, True) for r in range(world)]
