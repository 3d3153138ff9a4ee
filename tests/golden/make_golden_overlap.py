"""Goldens of the overlap-graph stage: the unmodified reference AdjList (oracle/_ref/AdjList-ref, built by
`make -C oracle ref`) run on the seeded contig sets of tests/overlap_cases.py.  Writes overlap_cases.json
(sha256 of the output of every case) and the complete output of three cases.

    python tests/golden/make_golden_overlap.py
"""
import hashlib
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import overlap_cases as oc  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "AdjList-ref")
FULL = {"unitigs_k64_adj", "unitigs_k32_dot", "fuzz7"}


def main():
    tmp = "/tmp/abyss_golden_overlap"
    os.makedirs(tmp, exist_ok=True)
    out = {}
    for c in oc.all_cases():
        fa = os.path.join(tmp, c["name"] + ".fa")
        oc.write_fasta(c, fa)
        r = subprocess.run([REF] + oc.command_args(c, fa), capture_output=True, check=True)
        data = oc.normalise(r.stdout, REF).replace(fa.encode(), b"IN.fa")
        out[c["name"]] = dict(sha256=hashlib.sha256(data).hexdigest(), bytes=len(data), contigs=len(c["records"]))
        if c["name"] in FULL:
            with open(os.path.join(HERE, "overlap_" + c["name"] + ".txt"), "wb") as f:
                f.write(data)
    with open(os.path.join(HERE, "overlap_cases.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(len(out), "cases")


if __name__ == "__main__":
    main()
