"""Pass-1 throughput probe on one GPU with device-resident synthetic reads (not the bench)."""
import sys, time, ctypes as C
import torch
sys.path.insert(0, ".")
from abyss_b200 import capi
from abyss_b200.synth import ReadSet
from abyss_b200.synth_torch import TorchReadSet

n_reads = int(float(sys.argv[1])) if len(sys.argv) > 1 else 2_000_000
budget = int(float(sys.argv[2])) if len(sys.argv) > 2 else 8 << 30
k, H, L = 64, 4, 150
counters = int(round(budget / 1.125)); counters += (-counters) % 64
genome = int(float(sys.argv[4])) if len(sys.argv) > 4 else int(n_reads * L / 40)
rs = ReadSet(2, genome, n_reads, L, 0.005)
t = TorchReadSet(rs, "cuda")
t0 = time.time()
chunks = [t.ascii(s, min(rs.n, s + (1 << 20))).reshape(-1) for s in range(0, rs.n, 1 << 20)]
bases = torch.cat(chunks); del chunks
offs = torch.arange(rs.n + 1, dtype=torch.int64, device="cuda") * L
torch.cuda.synchronize()
print(f"generated {rs.n} reads in {time.time()-t0:.1f}s")
windows = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1 << 16, 1 << 17, 1 << 18, 1 << 19]
for window in windows:
    f = capi.Filter.counting(counters, H, k, 3)
    f.set_window(window)
    f.insert_reads_dev(bases.data_ptr(), offs.data_ptr(), min(rs.n, 100000), 0)  # warm-up
    f.clear(); f.stats(reset=True)
    torch.cuda.synchronize(); t0 = time.time()
    n = f.insert_reads_dev(bases.data_ptr(), offs.data_ptr(), rs.n, bases.numel())
    dt = time.time() - t0
    st = f.stats()
    print(f"window {window:8d}: {n/dt/1e9:6.3f} G kmers/s wall | hash {st.ms_hash:8.1f} ms insert {st.ms_insert:8.1f} ms "
          f"-> insert-only {n/st.ms_insert/1e6:6.3f} G/s | windows {st.windows} deferred {st.deferred} ({100*st.deferred/n:.2f}%) launches {st.launches} drains {st.drains} "
          f"drained {st.drained_slots} | map 2^{__import__('os').environ.get('ABB_MAP_LOG2', '25')}")
    f.close()
