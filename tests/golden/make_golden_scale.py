#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: goldens at the sizes SURVEY.md 8(d) names, from the UNMODIFIED reference compiled
in oracle/_ref (`make -C oracle ref`; -j1 is deterministic).  Too large to commit as FASTA, so the fixture is
scale_cases.json: per case the md5 of the FASTA, of the --read-log, the sha256 of the counting filter
(`abyss-bloom build -t counting`), unitig count and total bases.  Reads are regenerated from the seed by
abyss_b200.synth on the GPU box.

  cfg1_k{32,40,48,64,96}   SURVEY 8(d) config 1: 200 kbp genome seed 1, 40x -> 53 333 x 150 bp, -kK --kc=2 -b64M -H4
  cfg1_edge_k32            the same reads through synth.edge_mutate (N, lower-case ends, short reads)
  m1_k64                   1 M x 150 bp of a 5 Mbp genome (seed 7, 30x), -k64 --kc=3 -b1G -H4

Run in the build container only:  python tests/golden/make_golden_scale.py [case ...]"""
import hashlib
import json
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from abyss_b200.synth import ReadSet, edge_mutate  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref")
DBG = os.path.join(REF, "abyss-bloom-dbg-ref")
BLOOM = os.path.join(REF, "abyss-bloom-ref")
TMP = "/tmp/abyss_golden_scale"

CASES = [dict(name=f"cfg1_k{k}", seed=1, genome=200000, n_reads=53333, L=150, err=0.005, k=k, kc=2, b="64M", H=4, edge=False)
         for k in (32, 40, 48, 64, 96)]
CASES.append(dict(name="cfg1_edge_k32", seed=1, genome=200000, n_reads=53333, L=150, err=0.005, k=32, kc=2, b="64M", H=4, edge=True))
CASES.append(dict(name="m1_k64", seed=7, genome=5000000, n_reads=1000000, L=150, err=0.005, k=64, kc=3, b="1G", H=4, edge=False))


def counters_for_budget(b):  # bloom-dbg.cc:359-367
    mult = {"k": 1 << 10, "M": 1 << 20, "G": 1 << 30}
    x = int(b[:-1]) * mult[b[-1]] / 1.125
    r = int(x + 0.5)
    return r if r % 64 == 0 else r + 64 - r % 64


def md5_file(path):
    h = hashlib.md5()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 22), b""):
            h.update(blk)
    return h.hexdigest()


def write_reads(c, path):
    rs = ReadSet(c["seed"], c["genome"], c["n_reads"], c["L"], c["err"])
    if not c["edge"]:
        rs.write_fastq(path)
        return
    seqs = edge_mutate([a.tobytes().decode() for a in rs.ascii(0, rs.n)])
    with open(path, "w") as f:
        for i, s in enumerate(seqs):
            f.write(f"@{rs.read_id(i)}\n{s}\n+\n{'I' * len(s)}\n")


def main():
    os.makedirs(TMP, exist_ok=True)
    path = os.path.join(HERE, "scale_cases.json")
    done = {c["name"]: c for c in json.load(open(path))} if os.path.exists(path) else {}
    want = set(sys.argv[1:])
    for c in CASES:
        if want and c["name"] not in want:
            continue
        fq = os.path.join(TMP, c["name"] + ".fq")
        fa = os.path.join(TMP, c["name"] + ".fa")
        log = os.path.join(TMP, c["name"] + ".readlog.tsv")
        bf = os.path.join(TMP, c["name"] + ".bloom")
        write_reads(c, fq)
        t0 = time.time()
        cmd = f"ulimit -s 65536; {DBG} -k{c['k']} --kc={c['kc']} -b{c['b']} -H{c['H']} -j1 --read-log={log} {fq} > {fa}"
        subprocess.run(["bash", "-c", cmd], check=True, capture_output=True)
        c["ref_seconds_j1"] = round(time.time() - t0, 1)
        counters = counters_for_budget(c["b"])
        subprocess.run([BLOOM, "build", "-k", str(c["k"]), "-t", "counting", f"-b{counters}", f"-H{c['H']}", "-j1", bf, fq],
                       check=True, capture_output=True)
        blob = open(bf, "rb").read()
        tag = b"[HeaderEnd]\n"
        raw = blob[blob.index(tag) + len(tag):]
        assert len(raw) == counters
        seqs = [l.strip() for l in open(fa) if not l.startswith(">")]
        c.update(counters=counters, counters_sha256=hashlib.sha256(raw).hexdigest(), fasta_md5=md5_file(fa),
                 readlog_md5=md5_file(log), n_contigs=len(seqs), bases=sum(map(len, seqs)))
        os.remove(bf)
        done[c["name"]] = c
        print(c, flush=True)
        json.dump([done[k] for k in sorted(done)], open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
