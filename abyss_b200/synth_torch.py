"""torch (CPU or CUDA) twin of abyss_b200.synth.ReadSet: bit-identical reads, generated where they
are needed (on the GPU for the device-resident bench leg, so 50M reads take seconds, not minutes).
Only plumbing: torch is used for device memory and integer tensor ops, nothing on the hot path."""
from __future__ import annotations

import torch

from .synth import ReadSet, _stream

_M64 = (1 << 64) - 1


def _s64(x: int) -> int:
    """python int (uint64 bit pattern) -> signed int64 value with the same bits"""
    x &= _M64
    return x - (1 << 64) if x >= (1 << 63) else x


def _lsr(x: torch.Tensor, n: int) -> torch.Tensor:
    """logical shift right of int64 bit patterns"""
    return (x >> n) & ((1 << (64 - n)) - 1)


def _mix(seed: int, idx: torch.Tensor) -> torch.Tensor:
    z = idx * _s64(0x9E3779B97F4A7C15) + _s64(seed)
    z = (z ^ _lsr(z, 30)) * _s64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _s64(0x94D049BB133111EB)
    return z ^ _lsr(z, 31)


def _umod(x: torch.Tensor, m: int) -> torch.Tensor:
    """(uint64 bit pattern x) mod m for m < 2^31, via two halves"""
    hi = _lsr(x, 32)
    lo = x & 0xFFFFFFFF
    return (((hi % m) * ((1 << 32) % m)) % m + lo % m) % m


class TorchReadSet:
    def __init__(self, rs: ReadSet, device="cuda"):
        assert not rs.paired or rs.G < (1 << 31)
        assert rs.G < (1 << 31), "torch generator supports genomes < 2^31 bases"
        self.rs, self.device = rs, torch.device(device)
        self._genome = None

    @property
    def genome(self) -> torch.Tensor:
        if self._genome is None:
            G = self.rs.G
            out = torch.empty(G, dtype=torch.uint8, device=self.device)
            step = 1 << 26
            s0 = _stream(self.rs.seed, 0)
            for s in range(0, G, step):
                e = min(G, s + step)
                idx = torch.arange(s, e, dtype=torch.int64, device=self.device)
                out[s:e] = (_mix(s0, idx) & 3).to(torch.uint8)
            self._genome = out
        return self._genome

    def codes(self, lo: int, hi: int) -> torch.Tensor:
        rs = self.rs
        hi = min(hi, rs.n)
        L, G = rs.L, rs.G
        dev = self.device
        r = torch.arange(lo, hi, dtype=torch.int64, device=dev)
        s_pos, s_str, s_err = (_stream(rs.seed, i) for i in (1, 2, 3))
        if rs.paired:
            frag = r >> 1
            mate = (r & 1).bool()
            flen = 300 + _umod(_mix(_stream(rs.seed, 4), frag), 201)
            # modulus differs per element: (x mod (G - flen + 1)); do it in float-free integer form
            span = G - flen + 1
            x = _mix(s_pos, frag)
            hi32, lo32 = _lsr(x, 32), x & 0xFFFFFFFF
            fstart = (((hi32 % span) * ((1 << 32) % span)) % span + lo32 % span) % span
            flip = (_mix(s_str, frag) & 1).bool()
            second = mate ^ flip
            start = torch.where(second, fstart + flen - L, fstart)
            rev = second
        else:
            start = _umod(_mix(s_pos, r), G - L + 1)
            rev = (_mix(s_str, r) & 1).bool()
        cols = torch.arange(L, dtype=torch.int64, device=dev)
        idx = start[:, None] + torch.where(rev[:, None], (L - 1) - cols[None, :], cols[None, :])
        c = self.genome[idx]
        c = torch.where(rev[:, None], 3 - c, c)
        if rs.err > 0:
            e_idx = r[:, None] * L + cols[None, :]
            u = _mix(s_err, e_idx)
            hit = _lsr(u, 11) < int(rs._err_thresh)
            delta = ((u & 0x7FF) % 3 + 1).to(torch.uint8)
            c = torch.where(hit, (c + delta) & 3, c)
        return c.to(torch.uint8)

    def ascii(self, lo: int, hi: int) -> torch.Tensor:
        lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=self.device)
        return lut[self.codes(lo, hi).long()]
