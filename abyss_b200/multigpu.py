"""Multi-GPU pass 1: hash-range sharding of the k-mers + NCCL union.

One process per GPU (torch.distributed, backend nccl; gloo for the CPU tests of the host logic).
Reads are split into contiguous file-order slices, one per rank.  Every rank hashes its slice (K1),
routes each canonical hash to the rank that owns its hash range with ONE all-to-all, inserts the
k-mers it owns -- the receive buffer concatenated in source-rank order IS file order -- into its
full-size private counting filter with the ordered insert, and the filters are merged with an
all-reduce(max) over NVLink (NCCL has no bitwise OR; for bit filters max over bytes of disjoint...
is not OR, so bit filters use all-gather + OR instead -- not needed by this path).

Exactness: every occurrence of a k-mer lands on one rank in file order, so each k-mer's own count is
what the sequential insert gives on that rank's filter; counters shared by k-mers of different
owners take the maximum instead of the conservative-update interplay.  The merged filter therefore
never under-counts a k-mer (minCount >= its exact multiplicity, capped at 255) and has no more false
positives than the reference's; unitigs agree with the single-GPU run as a set, not byte-for-byte
in their coverage figures.  torch is plumbing only (device memory, collectives).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def owner_of(h0: torch.Tensor, world: int) -> torch.Tensor:
    """contiguous hash ranges: owner = floor(top16(h0) * world / 65536); h0 is an int64 view of uint64"""
    top = (h0 >> 48) & 0xFFFF
    return (top * world) >> 16


def route_by_owner(h0: torch.Tensor, valid: torch.Tensor, world: int):
    """stable partition of the valid hashes by owner: returns (send buffer, per-destination counts)"""
    # destination of every slot (world for invalid slots); one order-preserving selection per destination
    # (torch.sort is limited to 2^31 elements; a rank's slice of the 50 M-read workload has more)
    own = owner_of(h0, world).to(torch.uint8)
    own[~valid.bool()] = world
    counts = torch.bincount(own.to(torch.int32), minlength=world + 1)[:world].to(torch.int64)
    send = torch.empty(int(counts.sum()), dtype=h0.dtype, device=h0.device)  # filled destination by destination
    at = 0
    for g in range(world):
        n = int(counts[g])
        torch.masked_select(h0, own == g, out=send[at:at + n])
        at += n
    return send, counts


def exchange(send: torch.Tensor, counts: torch.Tensor, group=None) -> torch.Tensor:
    """all-to-all of variable-size chunks; the result is ordered by source rank"""
    world = dist.get_world_size(group)
    recv_counts = torch.empty_like(counts)
    dist.all_to_all_single(recv_counts, counts, group=group)
    in_split = [int(x) for x in counts.tolist()]
    out_split = [int(x) for x in recv_counts.tolist()]
    recv = torch.empty(sum(out_split), dtype=send.dtype, device=send.device)
    dist.all_to_all_single(recv, send, output_split_sizes=out_split, input_split_sizes=in_split, group=group)
    assert len(in_split) == world
    return recv


class _DevArray:
    """expose a raw device pointer to torch through __cuda_array_interface__"""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def filter_tensor(filt, device) -> torch.Tensor:
    return torch.as_tensor(_DevArray(filt.device_ptr(), filt.sizeInBytes()), device=device)


def sharded_insert(filt, bases: torch.Tensor, offs: torch.Tensor, n_reads: int, group=None) -> int:
    """pass 1 over this rank's slice of the reads; on return every rank holds the merged filter.
    Returns the number of k-mers this rank inserted (owned)."""
    world = dist.get_world_size(group)
    dev = bases.device
    slots = filt.hash_reads_dev(bases.data_ptr(), offs.data_ptr(), n_reads)
    h0 = torch.empty(max(slots, 1), dtype=torch.int64, device=dev)
    valid = torch.empty(max(slots, 1), dtype=torch.uint8, device=dev)
    if slots:
        filt.hash_reads_dev(bases.data_ptr(), offs.data_ptr(), n_reads, h0.data_ptr(), valid.data_ptr(), slots)
    # stable partition by owner with the library's selection kernels (route_by_owner is the torch statement
    # of the same thing, kept for the CPU tests of the host logic)
    send = torch.empty(max(slots, 1), dtype=torch.int64, device=dev)
    cnt = filt.route_h0_dev(h0.data_ptr(), valid.data_ptr(), slots, world, send.data_ptr())
    counts = torch.from_numpy(cnt.astype("int64")).to(dev)
    send = send[:int(cnt.sum())]
    del h0, valid
    torch.cuda.empty_cache()
    recv = exchange(send, counts, group)
    del send
    torch.cuda.synchronize(dev)
    if recv.numel():
        filt.insert_h0_dev(recv.data_ptr(), recv.numel())
    owned = recv.numel()
    del recv
    if dev.type == "cuda":
        torch.cuda.empty_cache()  # hand the routing buffers back: the library allocates with cudaMalloc
    t = filter_tensor(filt, dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    torch.cuda.synchronize(dev)
    return owned


def sharded_classify(asm, bases_slice: torch.Tensor, offs_slice: torch.Tensor, n_slice: int, n_total: int, group=None) -> torch.Tensor:
    """K3a on this rank's slice of the reads, all-gather of the per-read codes; the returned tensor holds the
    codes of all n_total reads (file order) and has been handed to `asm` for its next process_reads call.
    Slices are the contiguous ranges [r * n // world, (r + 1) * n // world)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = bases_slice.device
    sizes = [(r + 1) * n_total // world - r * n_total // world for r in range(world)]
    assert sizes[rank] == n_slice
    mx = max(sizes)
    mine = torch.zeros(mx, dtype=torch.uint8, device=dev)
    if n_slice:
        asm.classify_dev(bases_slice.data_ptr(), offs_slice.data_ptr(), n_slice, mine.data_ptr())
    gathered = torch.empty(world * mx, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(gathered, mine, group=group)
    codes = torch.cat([gathered[r * mx:r * mx + sizes[r]] for r in range(world)])
    asm.set_codes(codes.data_ptr(), n_total)
    return codes  # keep alive until process_reads has run
