"""2-GPU test of the hash-range sharded pass 1 (run under `gpurun --gpus 2`; skipped with < 2 devices)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
from abyss_b200 import capi, multigpu
from abyss_b200.synth import ReadSet
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
dev = torch.device("cuda", rank)
rs = ReadSet.from_coverage(21, 20000, 30, 150, 0.005)
k, kc, H, m, L = 32, 2, 4, 932096, 150
asc = rs.ascii(0, rs.n)
lo, up = rank * rs.n // world, (rank + 1) * rs.n // world
bases = torch.from_numpy(asc[lo:up].reshape(-1).copy()).to(dev)
offs = torch.arange(up - lo + 1, dtype=torch.int64, device=dev) * L
f = capi.Filter.counting(m, H, k, kc, device=rank)
owned = multigpu.sharded_insert(f, bases, offs, up - lo)
merged = f.download()
# sharded classification: codes gathered from both ranks == codes of an unsharded run on the merged filter
all_bases = torch.from_numpy(asc.reshape(-1).copy()).to(dev)
all_offs = torch.arange(rs.n + 1, dtype=torch.int64, device=dev) * L
a1 = capi.Assembler(f)
codes = multigpu.sharded_classify(a1, bases, offs, up - lo, rs.n)
out_sharded = a1.process_reads_dev(all_bases.data_ptr(), all_offs.data_ptr(), rs.n)
res_sharded = a1.read_results().copy()
a1.close()
a2 = capi.Assembler(f)
out_plain = a2.process_reads_dev(all_bases.data_ptr(), all_offs.data_ptr(), rs.n)
res_plain = a2.read_results().copy()
a2.close()
assert out_sharded == out_plain and (res_sharded == res_plain).all()
tot = torch.tensor([owned], device=dev); dist.all_reduce(tot)
assert int(tot.item()) == rs.n * (L - k + 1), (int(tot.item()), rs.n * (L - k + 1))
# every rank holds the same merged filter
chk = torch.tensor([int(merged.astype(np.uint64).sum())], device=dev)
mx = chk.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX); assert int(mx.item()) == int(chk.item())
if rank == 0:
    # single-GPU filter of the same reads: the merged one never under-counts an inserted k-mer
    g = capi.Filter.counting(m, H, k, kc, device=0)
    g.insert_reads(capi.fixed_length_reads(asc))
    h0, valid, _ = capi.hash_reads(k, capi.fixed_length_reads(asc[:400]))
    mult = [np.uint64((i ^ ((k * 0x90b45d39fb6da1fa) & 0xFFFFFFFFFFFFFFFF)) & 0xFFFFFFFFFFFFFFFF) for i in range(H)]
    hh = [h0]
    with np.errstate(over="ignore"):
        for j in range(1, H):
            t = h0 * mult[j]; hh.append(t ^ (t >> np.uint64(27)))
    hs = np.stack(hh, axis=1)
    uniq, cnt = np.unique(h0, return_counts=True)
    assert (f.minCount(hs) >= 1).all()
    # unitigs from the merged filter equal the single-GPU ones as a set of canonical sequences
    ids = [rs.read_id(i) for i in range(rs.n)]
    def unitigs(filt):
        a = capi.Assembler(filt); out = a.process_reads(capi.fixed_length_reads(asc)); a.close()
        rc = lambda s: s.translate(str.maketrans("ACGT", "TGCA"))[::-1]
        return sorted(min(s, rc(s)) for _, s, _ in out)
    assert unitigs(f) == unitigs(g)
    print("MULTI_OK")
dist.barrier(); dist.destroy_process_group()
'''


def test_sharded_pass1_two_gpus(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    w = tmp_path / "worker.py"
    w.write_text(WORKER % {"root": ROOT})
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", str(w)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTI_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
