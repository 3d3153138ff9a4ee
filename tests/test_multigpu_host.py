"""CPU (gloo, world_size 2, 3 and 4): the host logic of the hash-range sharded pass 1 -- ownership, stable
routing, the all-to-all, and the file-order property of the receive buffer.  No GPU needed."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from abyss_b200.multigpu import exchange, owner_of, route_by_owner
    # global stream of "k-mers" in file order: value i carries its own index in the low bits
    n = 10000
    rng = np.random.default_rng(1)
    hi = rng.integers(0, 1 << 16, size=n, dtype=np.int64)
    allh = torch.from_numpy((hi << 48) | np.arange(n, dtype=np.int64))
    allv = torch.from_numpy((rng.random(n) > 0.1).astype(np.uint8))
    lo, up = rank * n // world, (rank + 1) * n // world  # contiguous file-order slices
    send, counts = route_by_owner(allh[lo:up], allv[lo:up], world)
    assert int(counts.sum()) == int(allv[lo:up].sum())
    recv = exchange(send, counts)
    # everything received is owned by this rank, valid, and in global file order
    assert bool((owner_of(recv, world) == rank).all())
    idx = recv & 0xFFFFFFFF
    assert bool((idx[1:] > idx[:-1]).all())
    expect = allh[(owner_of(allh, world) == rank) & allv.bool()]
    assert torch.equal(recv, expect)
    out[rank] = int(recv.numel())
    dist.barrier()
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("world", [2, 3, 4])
def test_sharded_routing(world):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert sum(out.values()) > 0 and len(out) == world


def test_owner_ranges():
    from abyss_b200.multigpu import owner_of
    h = torch.tensor([0, (1 << 62), -1, -(1 << 63)], dtype=torch.int64)  # int64 views of uint64 0, 2^62, 2^64-1, 2^63
    for world in (1, 2, 4, 8):
        o = owner_of(h, world)
        assert o.min() >= 0 and o.max() < world
        assert int(o[0]) == 0 and int(o[2]) == world - 1
