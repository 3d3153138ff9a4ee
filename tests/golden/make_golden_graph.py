"""Golden of the GraphViz dump (abyss-bloom-dbg -g FILE; outputGraph, bloom-dbg.h:1171-1242): the unmodified reference on a small
seeded read set (branches from sequencing errors, both strands).  The dump of the 2 kbp case is committed gzipped, of the
e2e_g10k_k25_small case as size + sha256.

    python tests/golden/make_golden_graph.py
"""
import gzip
import hashlib
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from abyss_b200.synth import ReadSet, edge_mutate  # noqa: E402

DBG = os.path.join(ROOT, "oracle", "_ref", "abyss-bloom-dbg-ref")
CASES = [
    dict(name="graph_g2k_k21", seed=31, genome=2000, cov=20, L=80, err=0.01, k=21, kc=2, b="64k", H=3),
    dict(name="graph_g10k_k25", seed=23, genome=10000, cov=25, L=100, err=0.01, k=25, kc=2, b="1M", H=3),
    # reads with 'N' (the trimmed read is the longest run of solid k-mers, a gap in the k-mer positions ends a run), lower-case
    # ends (removed by the reader) and reads shorter than k
    dict(name="graph_edge_k21", seed=33, genome=3000, cov=25, L=80, err=0.01, k=21, kc=2, b="256k", H=3, edge=True),
]


def write_reads(c, path):
    """the reads of a case as FASTQ (shared with the tests)"""
    rs = ReadSet.from_coverage(c["seed"], c["genome"], c["cov"], c["L"], c["err"])
    if not c.get("edge"):
        rs.write_fastq(path)
        return
    seqs = edge_mutate([a.tobytes().decode() for a in rs.ascii(0, rs.n)], every_n=5, every_lc=7, every_short=11)
    with open(path, "w") as f:
        for i, s in enumerate(seqs):
            f.write(f"@{rs.read_id(i)}\n{s}\n+\n{'I' * len(s)}\n")


def counters_for_budget(b):
    mult = {"k": 1 << 10, "M": 1 << 20, "G": 1 << 30}
    v = float(b[:-1]) * mult[b[-1]] if b[-1] in mult else float(b)
    c = int(round(v / 1.125))
    return c + (-c) % 64


def main():
    tmp = "/tmp/abyss_golden_graph"
    os.makedirs(tmp, exist_ok=True)
    out = []
    for c in CASES:
        fq = os.path.join(tmp, c["name"] + ".fq")
        write_reads(c, fq)
        dot = os.path.join(tmp, c["name"] + ".dot")
        cmd = f"ulimit -s 65536; {DBG} -k{c['k']} --kc={c['kc']} -b{c['b']} -H{c['H']} -j1 -g {dot} {fq} > /dev/null"
        subprocess.run(["bash", "-c", cmd], check=True)
        data = open(dot, "rb").read()
        c = dict(c, counters=counters_for_budget(c["b"]), bytes=len(data), sha256=hashlib.sha256(data).hexdigest(), lines=data.count(b"\n"))
        out.append(c)
        if c["name"] == "graph_g2k_k21":
            with gzip.GzipFile(os.path.join(HERE, c["name"] + ".dot.gz"), "wb", mtime=0) as f:
                f.write(data)
    json.dump(out, open(os.path.join(HERE, "graph_cases.json"), "w"), indent=1)
    print(out)


if __name__ == "__main__":
    main()
