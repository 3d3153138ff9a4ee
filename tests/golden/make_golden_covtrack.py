"""Golden of the 0/1 k-mer coverage track (abyss-bloom-dbg -C FILE -R REF; writeCovTrack, bloom-dbg.h:1280-1334): the
unmodified reference run on the reads of e2e_g20k_k32 with a reference FASTA made of the true genome (multi-line, lower-case
stretch), a copy with substitutions and 'N's, a record shorter than k and an unrelated sequence.

    python tests/golden/make_golden_covtrack.py
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from abyss_b200.synth import ReadSet  # noqa: E402

DBG = os.path.join(ROOT, "oracle", "_ref", "abyss-bloom-dbg-ref")


def ref_fasta(rs, path):
    """deterministic from the read set's seed: the GPU test rebuilds the same file"""
    g = np.frombuffer(b"ACGT", dtype=np.uint8)[rs.genome].tobytes().decode()
    rng = np.random.default_rng(rs.seed)
    mut = list(g[:6000])
    for p in rng.integers(0, len(mut), 60):
        mut[p] = "ACGT"[(("ACGT".index(mut[p]) if mut[p] in "ACGT" else 0) + 1) % 4]
    for p in rng.integers(0, len(mut), 8):
        mut[p] = "N"
    other = "".join("ACGT"[i] for i in rng.integers(0, 4, 3000))
    with open(path, "w") as f:
        f.write(">chrTrue the genome, 70 columns\n")
        body = g[:5000] + g[5000:5300].lower() + g[5300:]
        for i in range(0, len(body), 70):
            f.write(body[i:i + 70] + "\n")
        f.write(">chrMut substitutions and N\n" + "".join(mut) + "\n")
        f.write(">tiny\nACGTACGTAC\n")
        f.write(">unrelated\n" + other + "\n")


def main():
    import json
    c = {c["name"]: c for c in json.load(open(os.path.join(HERE, "e2e_cases.json")))}["e2e_g20k_k32"]
    rs = ReadSet.from_coverage(c["seed"], c["genome"], c["cov"], c["L"], c["err"])
    tmp = "/tmp/abyss_golden_cov"
    os.makedirs(tmp, exist_ok=True)
    fq = os.path.join(tmp, "r.fq")
    rs.write_fastq(fq)
    ref = os.path.join(tmp, "ref.fa")
    ref_fasta(rs, ref)
    out = os.path.join(HERE, "covtrack_g20k_k32.wig")
    cmd = f"ulimit -s 65536; {DBG} -k{c['k']} --kc={c['kc']} -b{c['b']} -H{c['H']} -j1 -C {out} -R {ref} {fq} > /dev/null"
    subprocess.run(["bash", "-c", cmd], check=True)
    print(out, sum(1 for _ in open(out)), "lines")


if __name__ == "__main__":
    main()
