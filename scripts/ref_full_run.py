#!/usr/bin/env python
"""One FULL run of the BASELINE config (50 M x 150 bp, k=64 kc=3 -b8G) through the unmodified reference with all host
threads, next to the same FASTQ through this repo's abyss-bloom-dbg (C++ CLI over libabyssb200): wall times, FASTA md5
and an order/strand-independent digest of the unitig sequences.  The reference at -j>1 is not deterministic (its counters
depend on the thread interleaving, SURVEY.md section 7 hard part 1), so byte equality is only expected against -j1 (the
tests); here the two unitig SETS are compared.   python scripts/ref_full_run.py [n_reads] > gpurun_out/ref_full_run.json"""
import hashlib
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from abyss_b200.synth import ReadSet  # noqa: E402


def digest(path):
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    md5, canon, lens = hashlib.md5(), [], []
    with open(path, "rb") as f:
        for line in f:
            md5.update(line)
            if not line.startswith(b">"):
                s = line.strip()
                canon.append(hashlib.md5(min(s, s.translate(comp)[::-1])).digest())
                lens.append(len(s))
    return {"fasta_md5": md5.hexdigest(), "unitig_multiset_md5": hashlib.md5(b"".join(sorted(canon))).hexdigest(),
            "unitigs": len(lens), "bases": sum(lens)}, set(canon)


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else bench.N_READS
    rs = ReadSet(bench.SEED, bench.GENOME, n, bench.L, bench.ERR, paired=True)
    d = "/dev/shm/abyss_full"
    os.makedirs(d, exist_ok=True)
    fq = os.path.join(d, "reads.fq")
    t0 = time.time()
    bench.write_sample_fastq(rs, n, fq)
    t_gen = time.time() - t0
    cores = os.cpu_count()
    ref_fa, our_fa = os.path.join(d, "ref.fa"), os.path.join(d, "ours.fa")
    t_ref = bench.run_reference(fq, cores, ref_fa)
    exe = os.path.join(ROOT, "abyss_b200", "lib", "abyss-bloom-dbg")
    t0 = time.time()
    r = subprocess.run([exe, f"-k{bench.K}", f"--kc={bench.KC}", f"-b{bench.BLOOM_BYTES}", f"-H{bench.H}", f"-j{min(cores, 16)}", "-o", our_fa, fq],
                       capture_output=True, text=True)
    t_ours = time.time() - t0
    assert r.returncode == 0, r.stderr[-2000:]
    dr, sr = digest(ref_fa)
    do, so = digest(our_fa)
    kmers = n * (bench.L - bench.K + 1)
    out = {"reads": n, "kmers": kmers, "fastq_bytes": os.path.getsize(fq), "fastq_generation_s": t_gen, "host_cores": cores,
           "reference": dict(dr, wall_s=t_ref, kmers_per_s=kmers / t_ref, cmd=f"abyss-bloom-dbg-ref -k64 --kc=3 -b8G -H4 -j{cores}"),
           "ours_cli": dict(do, wall_s=t_ours, kmers_per_s=kmers / t_ours,
                            cmd="abyss_b200/lib/abyss-bloom-dbg (same options; wall time includes process start, FASTQ parsing on the host, both passes)"),
           "unitigs_in_both": len(sr & so), "only_reference": len(sr - so), "only_ours": len(so - sr),
           "speedup_wall": t_ref / t_ours}
    print(json.dumps(out, indent=1))
    for f in (fq, ref_fa, our_fa):
        os.remove(f)


if __name__ == "__main__":
    main()
