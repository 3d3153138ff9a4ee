#!/usr/bin/env python3
"""TEST INFRASTRUCTURE: checkpoint files of the UNMODIFIED reference (BloomDBG/Checkpoint.h) for e2e_g20k_k32 with
--checkpoint=1500 --keep-checkpoint -j1: sha256 of PREFIX.dbg.bloom / .visited.bloom / .contigs.fa and the text of
.counters.tsv (state after the last checkpoint, 3000 of 4000 reads) -> tests/golden/checkpoint_case.json."""
import hashlib
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from abyss_b200.synth import ReadSet  # noqa: E402

DBG = os.path.join(ROOT, "oracle", "_ref", "abyss-bloom-dbg-ref")
TMP = "/tmp/abyss_golden_ckpt"


def main():
    os.makedirs(TMP, exist_ok=True)
    for f in os.listdir(TMP):
        os.remove(os.path.join(TMP, f))
    c = {x["name"]: x for x in json.load(open(os.path.join(HERE, "e2e_cases.json")))}["e2e_g20k_k32"]
    rs = ReadSet.from_coverage(c["seed"], c["genome"], c["cov"], c["L"], c["err"])
    fq = os.path.join(TMP, "r.fq")
    rs.write_fastq(fq)
    pfx = os.path.join(TMP, "ck")
    cmd = f"ulimit -s 65536; {DBG} -k{c['k']} --kc={c['kc']} -b{c['b']} -H{c['H']} -j1 --checkpoint=1500 --keep-checkpoint --checkpoint-prefix={pfx} {fq} > {TMP}/out.fa"
    subprocess.run(["bash", "-c", cmd], check=True)
    sha = lambda p: hashlib.sha256(open(p, "rb").read()).hexdigest()
    out = {"case": "e2e_g20k_k32", "reads_per_checkpoint": 1500, "counters_tsv": open(pfx + ".counters.tsv").read(),
           "dbg_bloom_sha256": sha(pfx + ".dbg.bloom"), "visited_bloom_sha256": sha(pfx + ".visited.bloom"),
           "contigs_fa_sha256": sha(pfx + ".contigs.fa")}
    json.dump(out, open(os.path.join(HERE, "checkpoint_case.json"), "w"), indent=1)
    print(out)


if __name__ == "__main__":
    main()
