"""CPU (gloo, world_size 1, 2, 3 and 4): the protocol of the position-sharded exact insert -- partial minima and veto
flags, ONE all-reduce(min) per step, identical file-order carry lists on every rank -- reproduces the sequential
oracle's counters bit for bit at every world size.  The CUDA implementation (abb_shard.cuh) follows the same steps;
tests/shard_model.py is its numpy statement.  No GPU needed."""
import os
import socket

import numpy as np
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


def _case(seed, n_reads, L, k, H, m):
    import oracle_py
    from abyss_b200.synth import ReadSet
    orc = oracle_py.load()
    rs = ReadSet(seed, 3000, n_reads, L, 0.01)
    seqs = [a.tobytes() for a in rs.ascii(0, rs.n)]
    seqs[3] = seqs[3][:20] + b"N" + seqs[3][21:]          # a window with a non-ACGT base yields no k-mer
    seqs += [b"T" * 60, b"T" * 60]                          # one k-mer repeated: long dependency chain
    hs = np.concatenate([orc.hash_seq(s, k, H)[0] for s in seqs])
    exp = np.zeros(m, dtype=np.uint8)
    orc.cbf_insert_hashes(exp, hs)
    return (hs % np.uint64(m)).astype(np.int64), exp


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from shard_model import shard_range, sharded_insert_model
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def allreduce_min(buf):
        t = torch.from_numpy(buf.copy())
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return t.numpy()

    ok = True
    # tiny filter (everything conflicts, saturation), a medium one, and one with H = 3
    for (seed, n_reads, L, k, H, m, window, ent, lanes) in [(5, 60, 60, 20, 4, 512, 64, 1 << 8, 16), (6, 120, 80, 25, 4, 40000, 256, 1 << 10, 32),
                                                             (7, 80, 60, 21, 3, 3001 * 8, 100, 1 << 9, 8)]:
        pos, exp = _case(seed, n_reads, L, k, H, m)
        got, steps = sharded_insert_model(pos, m, rank, world, allreduce_min, window, ent, lanes)
        lo, hi = shard_range(m, rank, world)
        ok &= bool((got[lo:hi] == exp[lo:hi]).all())
        # all-gather of the shards = the whole sequential array
        parts = [torch.zeros(m, dtype=torch.uint8) for _ in range(world)]
        mine = torch.zeros(m, dtype=torch.uint8)
        mine[lo:hi] = torch.from_numpy(got[lo:hi])
        if world > 1:
            dist.all_gather(parts, mine)
            full = torch.stack(parts).max(dim=0).values.numpy()
        else:
            full = mine.numpy()
        ok &= bool((full == exp).all())
    out[rank] = ok
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2, 3, 4])
def test_sharded_insert_protocol_is_exact(world):
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world and all(out.values())
