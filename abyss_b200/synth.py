"""Deterministic synthetic read generator.

Counter-based (splitmix64 of ``seed ^ stream``, index) so that any slice of the data set can
be produced independently, bit-identically, on any machine and in any chunk size -- the GPU
box regenerates exactly the reads the golden fixtures were made from.

Model: uniform-random ACGT genome of ``genome_len`` bases; ``n_reads`` reads of ``read_len``
bases; start uniform in [0, G-L]; strand reversed with p=0.5; each base substituted by one of
the three other bases with probability ``err``.  ``paired=True`` draws fragments of
300..500 bp and emits mate 1 / mate 2 (opposite strands) as reads 2i / 2i+1.
Quality strings are all 'I' (the reference stage ignores qualities unless -q/-Q is given).
"""
from __future__ import annotations

import numpy as np

_U64 = np.uint64
_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


def _mix(seed: int, idx: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser of (seed + idx * golden); idx is a uint64 array."""
    with np.errstate(over="ignore"):
        z = idx.astype(_U64) * _U64(0x9E3779B97F4A7C15) + _U64(seed & 0xFFFFFFFFFFFFFFFF)
        z = (z ^ (z >> _U64(30))) * _U64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> _U64(27))) * _U64(0x94D049BB133111EB)
        return z ^ (z >> _U64(31))


def _stream(seed: int, stream: int) -> int:
    return int(_mix(seed, np.array([stream + 1], dtype=_U64))[0])


def genome_codes(seed: int, start: int, n: int) -> np.ndarray:
    """2-bit codes (0..3 = A,C,G,T) of genome[start:start+n]."""
    idx = np.arange(start, start + n, dtype=_U64)
    return (_mix(_stream(seed, 0), idx) & _U64(3)).astype(np.uint8)


class ReadSet:
    """A reproducible synthetic read set; ``codes(lo, hi)`` / ``ascii(lo, hi)`` give reads lo..hi."""

    def __init__(self, seed: int, genome_len: int, n_reads: int, read_len: int = 150,
                 err: float = 0.005, paired: bool = False):
        assert genome_len >= (500 if paired else read_len)
        self.seed, self.G, self.n, self.L = seed, genome_len, n_reads, read_len
        self.err, self.paired = err, paired
        self._genome = None
        self._err_thresh = _U64(int(err * float(1 << 53)))

    @classmethod
    def from_coverage(cls, seed, genome_len, cov, read_len=150, err=0.005, paired=False):
        return cls(seed, genome_len, int(genome_len * cov / read_len), read_len, err, paired)

    @property
    def genome(self) -> np.ndarray:
        if self._genome is None:
            out = np.empty(self.G, dtype=np.uint8)
            step = 1 << 24
            for s in range(0, self.G, step):
                e = min(self.G, s + step)
                out[s:e] = genome_codes(self.seed, s, e - s)
            self._genome = out
        return self._genome

    def kmers_per_read(self, k: int) -> int:
        return max(0, self.L - k + 1)

    def codes(self, lo: int, hi: int) -> np.ndarray:
        """(hi-lo, L) uint8 array of 2-bit base codes for reads lo..hi-1."""
        lo, hi = int(lo), int(min(hi, self.n))
        L, G = self.L, self.G
        r = np.arange(lo, hi, dtype=_U64)
        s_pos, s_str, s_err = (_stream(self.seed, i) for i in (1, 2, 3))
        if self.paired:
            frag = r >> _U64(1)
            mate = (r & _U64(1)).astype(bool)
            flen = _U64(300) + _mix(_stream(self.seed, 4), frag) % _U64(201)
            fstart = _mix(s_pos, frag) % (_U64(G) - flen + _U64(1))
            flip = (_mix(s_str, frag) & _U64(1)).astype(bool)
            # mate 1 reads the fragment's 5' end forward, mate 2 its 3' end reverse-complemented;
            # a flipped fragment swaps the two roles
            second = mate ^ flip
            start = np.where(second, fstart + flen - _U64(L), fstart)
            rev = second
        else:
            start = _mix(s_pos, r) % _U64(G - L + 1)
            rev = (_mix(s_str, r) & _U64(1)).astype(bool)
        cols = np.arange(L, dtype=np.int64)
        idx = start.astype(np.int64)[:, None] + np.where(rev[:, None], (L - 1) - cols[None, :], cols[None, :])
        c = self.genome[idx]
        c = np.where(rev[:, None], 3 - c, c).astype(np.uint8)
        if self.err > 0:
            e_idx = (r[:, None] * _U64(L)) + cols[None, :].astype(_U64)
            u = _mix(s_err, e_idx)
            hit = (u >> _U64(11)) < self._err_thresh
            delta = (((u & _U64(0x7FF)) % _U64(3)) + _U64(1)).astype(np.uint8)
            c = np.where(hit, (c + delta) & 3, c).astype(np.uint8)
        return c

    def ascii(self, lo: int, hi: int) -> np.ndarray:
        """(hi-lo, L) uint8 array of ASCII bases."""
        return _BASES[self.codes(lo, hi)]

    def read_id(self, i: int) -> str:
        return f"r{i >> 1}/{(i & 1) + 1}" if self.paired else f"r{i}/1"

    def write_fastq(self, path: str, lo: int = 0, hi: int | None = None, fasta: bool = False,
                    chunk: int = 1 << 16) -> None:
        hi = self.n if hi is None else min(hi, self.n)
        qual = b"I" * self.L
        with open(path, "wb") as f:
            for s in range(lo, hi, chunk):
                e = min(hi, s + chunk)
                a = self.ascii(s, e)
                parts = []
                for j in range(e - s):
                    rid = self.read_id(s + j).encode()
                    if fasta:
                        parts.append(b">" + rid + b"\n" + a[j].tobytes() + b"\n")
                    else:
                        parts.append(b"@" + rid + b"\n" + a[j].tobytes() + b"\n+\n" + qual + b"\n")
                f.write(b"".join(parts))


def write_fastq_fast(rs: "ReadSet", path: str, lo: int, hi: int, chunk: int = 1 << 18, ascii_fn=None) -> None:
    """Fixed-width vectorised FASTQ writer for large bench samples (ids are zero-padded: @r0000012345/1, which only
    changes the `read:` label of the reference's FASTA headers).  ~20x faster than ReadSet.write_fastq."""
    hi = min(hi, rs.n)
    L = rs.L
    W = 10
    rec = 2 + W + 2 + 1 + L + 1 + 2 + L + 1  # "@r" id "/1" \n seq \n "+\n" qual \n
    with open(path, "wb") as f:
        for s in range(lo, hi, chunk):
            e = min(hi, s + chunk)
            n = e - s
            buf = np.empty((n, rec), dtype=np.uint8)
            idx = np.arange(s, e, dtype=np.int64)
            num = (idx >> 1) if rs.paired else idx
            buf[:, 0], buf[:, 1] = ord("@"), ord("r")
            for d in range(W):
                buf[:, 2 + W - 1 - d] = (num // (10 ** d)) % 10 + ord("0")
            buf[:, 2 + W] = ord("/")
            buf[:, 3 + W] = ((idx & 1) + 1 if rs.paired else 1) + ord("0")
            buf[:, 4 + W] = ord("\n")
            buf[:, 5 + W:5 + W + L] = ascii_fn(s, e) if ascii_fn else rs.ascii(s, e)
            o = 5 + W + L
            buf[:, o], buf[:, o + 1], buf[:, o + 2] = ord("\n"), ord("+"), ord("\n")
            buf[:, o + 3:o + 3 + L] = ord("I")
            buf[:, o + 3 + L] = ord("\n")
            f.write(buf.tobytes())


def revcomp(seq: str) -> str:
    return seq.translate(str.maketrans("ACGTacgt", "TGCAtgca"))[::-1]


def edge_mutate(seqs, every_n: int = 37, every_lc: int = 53, every_short: int = 211):
    """Deterministic edge-case variant of a read list (scale goldens, tests/golden/make_golden_scale.py):
    read i gets an 'N' at column (7*i) % len when i % every_n == 5, lower-case ends when i % every_lc == 11
    (15 bases each side: trimMasked removes them, FastaReader.cpp:29), is cut to 40 bases when
    i % every_short == 3."""
    out = []
    for i, s in enumerate(seqs):
        if i % every_n == 5 and s:
            c = (7 * i) % len(s)
            s = s[:c] + "N" + s[c + 1:]
        if i % every_lc == 11 and len(s) > 40:
            s = s[:15].lower() + s[15:-15] + s[-15:].lower()
        if i % every_short == 3:
            s = s[:40]
        out.append(s)
    return out
