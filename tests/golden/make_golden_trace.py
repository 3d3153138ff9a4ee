#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: `-T` trace goldens (ContigRecord rows, bloom-dbg.h:186-254) of the e2e cases from the UNMODIFIED
reference (oracle/_ref/abyss-bloom-dbg-ref -j1).  The `length` column of redundant rows is an uninitialised value in the
reference (ContigRecord() does not set it, outputContig only assigns it for printed contigs), so the generator blanks it."""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from abyss_b200.synth import ReadSet  # noqa: E402

DBG = os.path.join(ROOT, "oracle", "_ref", "abyss-bloom-dbg-ref")
TMP = "/tmp/abyss_golden"


def main():
    os.makedirs(TMP, exist_ok=True)
    for c in json.load(open(os.path.join(HERE, "e2e_cases.json"))):
        rs = ReadSet.from_coverage(c["seed"], c["genome"], c["cov"], c["L"], c["err"])
        fq = os.path.join(TMP, c["name"] + ".fq")
        rs.write_fastq(fq)
        tr = os.path.join(TMP, c["name"] + ".trace")
        cmd = f"ulimit -s 65536; {DBG} -k{c['k']} --kc={c['kc']} -b{c['b']} -H{c['H']} -j1 -T {tr} {fq} > /dev/null"
        subprocess.run(["bash", "-c", cmd], check=True)
        rows = [l.rstrip("\n").split("\t") for l in open(tr)]
        for r in rows[1:]:
            if r[2] == "1":
                r[1] = "-"
        with open(os.path.join(HERE, c["name"] + ".trace.tsv"), "w") as f:
            f.write("".join("\t".join(r) + "\n" for r in rows))
        print(c["name"], len(rows) - 1, "rows")


if __name__ == "__main__":
    main()
