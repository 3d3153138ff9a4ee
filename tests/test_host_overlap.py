"""CPU: the overlap-graph join logic (abyss_b200/csrc/abb_overlap.cuh -- the SAME per-item functions the CUDA kernels
call) and the product's AdjList command line and graph writers (abyss_b200/host/adjlist_main.h), run by the
single-thread harness tests/host_overlap, against the unmodified reference AdjList: committed goldens
(tests/golden/make_golden_overlap.py) and, where oracle/_ref/AdjList-ref exists, live on further seeds."""
import hashlib
import json
import os
import subprocess

import pytest

import overlap_cases as oc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
REF = os.path.join(ROOT, "oracle", "_ref", "AdjList-ref")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("ho") / "AdjList")
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wno-unknown-pragmas", "-o", exe, os.path.join(ROOT, "tests", "host_overlap", "host_overlap.cpp")],
                   check=True, capture_output=True)
    return exe


def run_case(exe, case, tmp_path):
    fa = str(tmp_path / (case["name"] + ".fa"))
    oc.write_fasta(case, fa)
    r = subprocess.run([exe] + oc.command_args(case, fa), capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    return oc.normalise(r.stdout, exe).replace(fa.encode(), b"IN.fa")


def test_goldens(harness, tmp_path):
    want = json.load(open(os.path.join(GOLD, "overlap_cases.json")))
    cases = oc.all_cases()
    assert sorted(c["name"] for c in cases) == sorted(want)
    for c in cases:
        got = run_case(harness, c, tmp_path)
        assert len(got) == want[c["name"]]["bytes"], c["name"]
        assert hashlib.sha256(got).hexdigest() == want[c["name"]]["sha256"], c["name"]
        full = os.path.join(GOLD, "overlap_" + c["name"] + ".txt")
        if os.path.exists(full):
            assert got == open(full, "rb").read()


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/AdjList-ref not built")
def test_live_against_reference(harness, tmp_path):
    for seed in range(1000, 1120):
        c = oc.fuzz_case(seed)
        fa = str(tmp_path / "in.fa")
        oc.write_fasta(c, fa)
        a = subprocess.run([REF] + oc.command_args(c, fa), capture_output=True)
        assert a.returncode == 0, a.stderr.decode()
        b = subprocess.run([harness] + oc.command_args(c, fa), capture_output=True)
        assert b.returncode == 0, b.stderr.decode()
        assert oc.normalise(a.stdout, REF) == oc.normalise(b.stdout, harness), (seed, c["k"], c["m"], c["args"])


def test_errors(harness, tmp_path):
    fa = str(tmp_path / "n.fa")
    open(fa, "w").write(">0 12 3\nACGTNACGTACG\n>1 12 3\nACGTACGTACGA\n")
    r = subprocess.run([harness, "-k6", fa], capture_output=True, text=True)
    assert r.returncode != 0 and "nucleotide" in r.stderr  # the reference's Kmer constructor aborts on the N
    open(fa, "w").write(">0\nACGT\n")
    r = subprocess.run([harness, "-k6", fa], capture_output=True, text=True)
    assert r.returncode != 0 and "not longer than k-1" in r.stderr
    open(fa, "w").write(">a\nACGTACGTAA\n>a\nACGTACGTAC\n")
    r = subprocess.run([harness, "-k6", fa], capture_output=True, text=True)
    assert r.returncode != 0 and "duplicate ID" in r.stderr
    r = subprocess.run([harness, fa], capture_output=True, text=True)
    assert r.returncode != 0 and "missing -k,--kmer option" in r.stderr


DBG_REF = os.path.join(ROOT, "oracle", "_ref", "abyss-bloom-dbg-ref")


@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(DBG_REF)), reason="oracle/_ref not built")
@pytest.mark.parametrize("k,m,fmt", [(32, 0, "--adj"), (32, 20, "--dot"), (48, 30, "--gfa1"), (64, 50, "--gfa2"), (96, 50, "--sam"), (40, 25, "--asqg")])
def test_real_unitig_sets(harness, tmp_path, k, m, fmt):
    # the pipeline of bin/abyss-pe on config 1 (53 333 x 150 bp reads of a 200 kbp genome): the reference's own
    # unitig FASTA (tips, branches, blunt ends from coverage gaps) into both AdjList implementations
    from abyss_b200.synth import ReadSet
    rs = ReadSet(1, 200000, 53333, 150, 0.005)
    fq = str(tmp_path / "r.fq")
    rs.write_fastq(fq)
    fa = str(tmp_path / "unitigs-1.fa")
    subprocess.run(["bash", "-c", f"ulimit -s 65536; {DBG_REF} -k{k} --kc=2 -b64M -H4 -j1 {fq} > {fa} 2>/dev/null"], check=True)
    n = sum(1 for line in open(fa) if line.startswith(">"))
    assert n > 10
    args = [f"-k{k}", f"-m{m}", fmt, fa]
    a = subprocess.run([REF] + args, capture_output=True)
    assert a.returncode == 0, a.stderr.decode()
    b = subprocess.run([harness] + args, capture_output=True)
    assert b.returncode == 0, b.stderr.decode()
    assert oc.normalise(a.stdout, REF) == oc.normalise(b.stdout, harness)
    assert len(a.stdout) > 0


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/AdjList-ref not built")
def test_option_aliases_and_stdin(harness, tmp_path):
    # --gv = --dot, --gfa = --gfa1, -m0 = k-1, long options, several input files, contigs on standard input
    c = oc.tiled_case(21, 20000, 31, 20)
    half = len(c["records"]) // 2
    a_fa, b_fa = str(tmp_path / "a.fa"), str(tmp_path / "b.fa")
    oc.write_fasta(dict(c, records=c["records"][:half]), a_fa)
    oc.write_fasta(dict(c, records=c["records"][half:]), b_fa)
    both = open(a_fa).read() + open(b_fa).read()
    for args, stdin in ((["--kmer=31", "--min-overlap=20", "--gv", a_fa, b_fa], None), (["-k31", "-m0", "--gfa", a_fa, b_fa], None),
                        (["-k", "31", "-m", "25", "--SS", "--adj"], both), (["-k31", "--no-SS", "--asqg", "-"], both)):
        ref = subprocess.run([REF] + args, input=stdin, capture_output=True, text=True)
        assert ref.returncode == 0, ref.stderr
        got = subprocess.run([harness] + args, input=stdin, capture_output=True, text=True)
        assert got.returncode == 0, got.stderr
        assert got.stdout == ref.stdout, args
