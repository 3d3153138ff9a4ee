"""Multi-GPU glue for Python harnesses (bench.py, tests).  One process per GPU.

Pass 1 is NOT here: the exact position-sharded insert and its NCCL communicator live behind the C ABI
(abb_comm_*, abb_insert_reads_sharded[_dev]; abyss_b200/csrc/abb_shard.cuh) and are reached through capi.Comm /
capi.Filter.insert_reads_sharded_dev.  torch.distributed is only used by the caller to ship the 128-byte NCCL id.

sharded_classify: K3a (processRead's per-read tests, bloom-dbg.h:803-817: pure per read) on this rank's slice of the
reads, all-gather of the per-read codes through the C-ABI communicator.  torch is device memory only.
"""
from __future__ import annotations

import torch


def slice_sizes(n_total: int, world: int):
    return [(r + 1) * n_total // world - r * n_total // world for r in range(world)]


def sharded_classify(asm, comm, d_bases_slice_ptr: int, d_offs_slice_ptr: int, n_slice: int, n_total: int, dev) -> torch.Tensor:
    """the returned tensor holds the codes of all n_total reads (file order) and has been handed to `asm` for its next
    process_reads call.  Slices are the contiguous ranges [r * n // world, (r + 1) * n // world)."""
    world, rank = comm.world, comm.rank
    sizes = slice_sizes(n_total, world)
    assert sizes[rank] == n_slice
    mx = (max(sizes) + 15) & ~15
    gathered = torch.zeros(world * mx, dtype=torch.uint8, device=dev)
    if n_slice:
        asm.classify_dev(d_bases_slice_ptr, d_offs_slice_ptr, n_slice, gathered.data_ptr() + rank * mx)
    torch.cuda.synchronize(dev)
    comm.allgather_bytes(gathered.data_ptr(), mx, 0)
    torch.cuda.synchronize(dev)
    codes = torch.cat([gathered[r * mx:r * mx + sizes[r]] for r in range(world)])
    asm.set_codes(codes.data_ptr(), n_total)
    return codes  # keep alive until process_reads has run
