"""Pass-1 probe of the position-sharded exact insert under torchrun (not the bench):
    torchrun --nproc-per-node N scripts/probe_sharded.py [n_reads] [genome]
Tuning knobs: ABB_SHARD_WINDOW (slots per global window), ABB_MAP_LOG2."""
import os, sys, time, hashlib
import torch, torch.distributed as dist
sys.path.insert(0, ".")
from abyss_b200 import capi
from abyss_b200.synth import ReadSet
from abyss_b200.synth_torch import TorchReadSet

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
dev = torch.device("cuda", rank)
n_reads = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4_000_000
genome = int(float(sys.argv[2])) if len(sys.argv) > 2 else 187_500_000
k, H, L = 64, 4, 150
counters = capi.counters_for_budget(8 << 30)
rs = ReadSet(2, genome, n_reads, L, 0.005)
t = TorchReadSet(rs, dev)
bases = torch.cat([t.ascii(s, min(rs.n, s + (1 << 20))).reshape(-1) for s in range(0, rs.n, 1 << 20)])
offs = torch.arange(rs.n + 1, dtype=torch.int64, device=dev) * L
def bcast(b):
    box = [b]; dist.broadcast_object_list(box, src=0); return box[0]
comm = capi.Comm(rank, world, rank, bcast)
f = capi.Filter.counting(counters, H, k, 3, device=rank)
f.insert_reads_sharded_dev(comm, bases.data_ptr(), offs.data_ptr(), min(rs.n, 100000))  # warm-up
f.clear(); f.stats(reset=True)
dist.barrier(); torch.cuda.synchronize(); t0 = time.time()
n = f.insert_reads_sharded_dev(comm, bases.data_ptr(), offs.data_ptr(), rs.n, finalize=False)
torch.cuda.synchronize(); dt = time.time() - t0
t1 = time.time(); f.allgather(comm); torch.cuda.synchronize(); tg = time.time() - t1
st = f.stats()
if rank == 0:
    print(f"world {world} window {os.environ.get('ABB_SHARD_WINDOW', 'default')} map 2^{os.environ.get('ABB_MAP_LOG2', '26')}: "
          f"{n/dt/1e9:6.3f} G kmers/s | hash {st.ms_hash:7.1f} ms insert {st.ms_insert:8.1f} ms allgather {tg*1e3:6.1f} ms | windows {st.windows} "
          f"deferred {100*st.deferred/n:.2f}% drains {st.drains} launches {st.launches}", flush=True)
dist.barrier(); dist.destroy_process_group()
