"""End-to-end probe (not the bench): pass 1 + pass 2 on one GPU, device-resident reads, with phase timings
and an md5 of the FASTA to compare against a reference run of the same seeded reads."""
import hashlib, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from abyss_b200 import capi
from abyss_b200.synth import ReadSet
from abyss_b200.synth_torch import TorchReadSet

seed, genome, cov = int(sys.argv[1]), int(float(sys.argv[2])), float(sys.argv[3])
k, kc, budget = int(sys.argv[4]), int(sys.argv[5]), int(float(sys.argv[6]))
want_md5 = sys.argv[7] if len(sys.argv) > 7 else None
H, L = 4, 150
rs = ReadSet.from_coverage(seed, genome, cov, L, 0.005)
t = TorchReadSet(rs, "cuda")
t0 = time.time()
bases = torch.cat([t.ascii(s, min(rs.n, s + (1 << 20))).reshape(-1) for s in range(0, rs.n, 1 << 20)])
offs = torch.arange(rs.n + 1, dtype=torch.int64, device="cuda") * L
torch.cuda.synchronize()
print(f"{rs.n} reads generated in {time.time()-t0:.1f}s")
counters = capi.counters_for_budget(budget)
f = capi.Filter.counting(counters, H, k, kc)
t0 = time.time()
nk = f.insert_reads_dev(bases.data_ptr(), offs.data_ptr(), rs.n, bases.numel())
t1 = time.time()
st = f.stats()
print(f"pass1: {nk} kmers in {t1-t0:.3f}s = {nk/(t1-t0)/1e9:.3f} G/s (hash {st.ms_hash:.0f} ms, insert {st.ms_insert:.0f} ms, deferred {st.deferred})")
print("popcount", f.popcounts(), "FPR(kc) %.3g" % f.filtered_FPR())
a = capi.Assembler(f)
t0 = time.time()
contigs = a.process_reads_dev(bases.data_ptr(), offs.data_ptr(), rs.n)
t2 = time.time()
s = a.stats(); c = a.counters()
print(f"pass2: {t2-t0:.3f}s -> {nk/(t2-t0)/1e9:.3f} G kmers/s | rounds {s.rounds} speculated {s.speculated_reads} wasted {s.wasted_reads} "
      f"candidates {s.candidates} contigs tried {s.contigs_tried} accepted {c.contig_id} bases {c.bases_assembled} launches {s.launches}")
print(f"       classify {s.ms_classify:.0f} ms visited {s.ms_visited:.0f} ms extend {s.ms_extend:.0f} ms replay {s.ms_replay:.0f} ms")
print(f"total: {nk/((t1-t0)+(t2-t0))/1e9:.3f} G kmers/s")
fa = "".join(f">{i} {len(sq)} {cv} read:{rs.read_id(sr)}\n{sq}\n" for i, (sr, sq, cv) in enumerate(contigs))
md5 = hashlib.md5(fa.encode()).hexdigest()
print("fasta md5", md5, "contigs", len(contigs), "MATCH" if want_md5 == md5 else ("MISMATCH" if want_md5 else ""))
