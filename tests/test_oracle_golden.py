"""The oracle (oracle/abyss_oracle.c) against the reference: the reference's own golden vector and
the committed fixtures produced by the UNMODIFIED reference (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from abyss_b200.synth import ReadSet

EDGE_FILE = "seqs_edge.txt"


def edge_seqs(golden_dir):
    return open(os.path.join(golden_dir, EDGE_FILE)).read().split("\n")[:-1]


def seeded(seed, genome, n, L, err=0.01):
    rs = ReadSet(seed, genome, n, L, err)
    return [a.tobytes().decode() for a in rs.ascii(0, n)]


def test_nthash_known_answer(oracle):
    # vendor/nthash/unittest/UnitTests.cpp:45-48
    h, pos = oracle.hash_seq("ACGTACACTGGACTGAGTCT", 20, 3)
    assert pos.tolist() == [0]
    assert h[0].tolist() == [10434435546371013747, 16073887395445158014, 8061578976118370557]


def test_bad_kmer_positions(oracle):
    # Unittest/BloomDBG/RollingHashIteratorTest.cpp:64-84: AAANAAA, k=3 -> positions 0 and 4
    h, pos = oracle.hash_seq("AAANAAA", 3, 2)
    assert pos.tolist() == [0, 4]
    assert (h[0] == h[1]).all()


def test_reverse_complement_invariance(oracle):
    # vendor/nthash/unittest/UnitTests.cpp:56-68
    s = "TGACTTTCGGGTGGAAAAGCTACGTACGTAAAGGGTTTCCCA"
    rc = s.translate(str.maketrans("ACGT", "TGCA"))[::-1]
    a, _ = oracle.hash_seq(s, 18, 3)
    b, _ = oracle.hash_seq(rc, 18, 3)
    assert (a == b[::-1]).all()


@pytest.mark.parametrize("name", ["hashes_k5", "hashes_k20", "hashes_k32", "hashes_k64", "hashes_mask", "hashes_mask33"])
def test_hash_fixtures(oracle, golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    k, H, mask = int(g["k"]), int(g["H"]), str(g["mask"])
    seqs = edge_seqs(golden_dir) + seeded(12, 20000, 600, 150)[:40]
    rows_seq, rows_pos, rows_h = [], [], []
    for i, s in enumerate(seqs):
        h, pos = oracle.hash_seq(s, k, H, mask)
        rows_seq += [i] * len(pos)
        rows_pos += pos.tolist()
        rows_h.append(h)
    assert rows_seq == g["seq"].tolist()
    assert rows_pos == g["pos"].tolist()
    assert (np.concatenate(rows_h) == g["h"]).all()


def test_counting_fixtures(oracle, golden_dir):
    reads60 = seeded(11, 3000, 1500, 60)
    reads150 = seeded(12, 20000, 600, 150)
    edge = edge_seqs(golden_dir)
    for name, seqs in (("count_m4096", edge + reads60), ("count_m65536_H3", reads150),
                       ("count_sat", [edge[9]] * 40 + reads60[:200])):
        g = np.load(os.path.join(golden_dir, name + ".npz"))
        c = np.zeros(int(g["m"]), dtype=np.uint8)
        oracle.cbf_load(c, seqs, int(g["k"]), int(g["H"]))
        assert (c == g["data"]).all(), name
    assert np.load(os.path.join(golden_dir, "count_sat.npz"))["data"].max() == 255  # saturation exercised


def test_bits_and_cascading_fixtures(oracle, golden_dir):
    reads60 = seeded(11, 3000, 1500, 60)
    edge = edge_seqs(golden_dir)
    g = np.load(os.path.join(golden_dir, "bits_m8192.npz"))
    b = np.zeros(int(g["m"]) // 8, dtype=np.uint8)
    oracle.bf_load(b, edge + reads60, int(g["k"]), int(g["H"]))
    assert (b == g["data"]).all()
    g = np.load(os.path.join(golden_dir, "casc_m8192_L3.npz"))
    L, m = int(g["L"]), int(g["m"])
    lv = np.zeros(L * m // 8, dtype=np.uint8)
    oracle.casc_load(lv, m, L, edge + reads60 + reads60[:700], int(g["k"]), int(g["H"]))
    assert (lv == g["data"]).all()


def test_counter_sizing(oracle):
    # bloom-dbg.cc:359-367 ; SURVEY.md section 8 sizes
    GiB = 1 << 30
    assert oracle.lib.abo_counters_for_budget(64 << 20) == 59652352
    assert oracle.lib.abo_counters_for_budget(GiB) == 954437184
    assert oracle.lib.abo_counters_for_budget(8 * GiB) == 7635497472
    assert oracle.lib.abo_counters_for_budget(64 * GiB) == 61083979328


def test_spaced_seed_known_answers():
    # Unittest/BloomDBG/SpacedSeedTest.cpp:16,25 (host-side mirror of SpacedSeed::qrSeed / qrSeedPair / kmerPair)
    from abyss_b200.capi import kmer_pair_seed, qr_seed_pair
    assert qr_seed_pair(22, 11)[:11] == "10100011101"
    assert qr_seed_pair(33, 11) == "101000111010000000000010111000101"
    assert kmer_pair_seed(10, 3) == "1110000111"
    with pytest.raises(ValueError):
        qr_seed_pair(33, 10)
    with pytest.raises(ValueError):
        kmer_pair_seed(10, 6)


def test_threshold_semantics_known_answer(oracle):
    # Unittest/BloomDBG/CountingBloomFilterTest.cpp:9-46: size 1000, H=1, threshold 2, the four 16-mers a..d;
    # a inserted twice and b once: only a passes the threshold
    a, b, c, d = "AGATGTGCTGCCGCCT", "TGGACAGCGTTACCTC", "TAATAACAGTCCCTAT", "GATCGTGGCGGGCGAT"
    counters = np.zeros(1000, dtype=np.uint8)
    oracle.cbf_load(counters, [a, a, b], 16, 1)
    mn = lambda s: int(oracle.cbf_min_hashes(counters, oracle.hash_seq(s, 16, 1)[0])[0])
    assert mn(a) == 2 and mn(b) == 1 and mn(c) == 0 and mn(d) == 0
    assert np.count_nonzero(counters) == 2 and np.count_nonzero(counters >= 2) == 1


def test_cascading_known_answer(oracle):
    # Unittest/BloomDBG/HashAgnosticCascadingBloomTest.cpp:9-46: 2 levels of 1000 bits... (size rounded to 1024 here),
    # H=1; a inserted twice, b once: the last level holds a only
    a, b, c = "AGATGTGCTGCCGCCT", "TGGACAGCGTTACCTC", "TAATAACAGTCCCTAT"
    mbits, L = 1024, 2
    levels = np.zeros(L * mbits // 8, dtype=np.uint8)
    oracle.casc_load(levels, mbits, L, [a, a, b], 16, 1)
    last = levels[(L - 1) * mbits // 8:]
    def in_last(s):
        p = int(oracle.hash_seq(s, 16, 1)[0][0][0] % np.uint64(mbits))
        return bool(last[p >> 3] >> (p & 7) & 1)
    assert in_last(a) and not in_last(b) and not in_last(c)
    first = levels[:mbits // 8]
    assert int(np.unpackbits(first).sum()) == 2 and int(np.unpackbits(last).sum()) == 1
