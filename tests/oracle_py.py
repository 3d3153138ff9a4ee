"""ctypes access to the C restatement oracle/abyss_oracle.c -- TEST INFRASTRUCTURE ONLY.
Nothing in abyss_b200/ imports this."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")
SO = os.path.join(ODIR, "_ref", "liboracle.so")

_vp = C.c_void_p
_lib = None


def build():
    src = [os.path.join(ODIR, f) for f in ("abyss_oracle.c", "abyss_oracle.h")]
    if not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in src):
        subprocess.run(["make", "-C", ODIR, "port"], check=True, capture_output=True)
    return SO


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        L = lib
        L.abo_hash_seq.restype = C.c_size_t
        L.abo_hash_seq.argtypes = [C.c_char_p, C.c_size_t, C.c_uint, C.c_uint, C.c_char_p, _vp, _vp]
        for fn in (L.abo_cbf_load_seq, L.abo_bf_load_seq):
            fn.restype = C.c_size_t
            fn.argtypes = [_vp, C.c_uint64, C.c_char_p, C.c_size_t, C.c_uint, C.c_uint, C.c_char_p]
        L.abo_casc_load_seq.restype = C.c_size_t
        L.abo_casc_load_seq.argtypes = [_vp, C.c_uint64, C.c_uint, C.c_char_p, C.c_size_t, C.c_uint, C.c_uint, C.c_char_p]
        L.abo_cbf_insert.restype = None
        L.abo_cbf_insert.argtypes = [_vp, C.c_uint64, _vp, C.c_uint]
        L.abo_cbf_min.restype = C.c_uint8
        L.abo_cbf_min.argtypes = [_vp, C.c_uint64, _vp, C.c_uint]
        L.abo_bf_insert.restype = None
        L.abo_bf_insert.argtypes = [_vp, C.c_uint64, _vp, C.c_uint]
        L.abo_bf_contains.restype = C.c_int
        L.abo_bf_contains.argtypes = [_vp, C.c_uint64, _vp, C.c_uint]
        L.abo_casc_insert.restype = None
        L.abo_casc_insert.argtypes = [_vp, C.c_uint64, C.c_uint, _vp, C.c_uint]
        L.abo_counters_for_budget.restype = C.c_uint64
        L.abo_counters_for_budget.argtypes = [C.c_uint64]
        L.abo_srol_n.restype = C.c_uint64
        L.abo_srol_n.argtypes = [C.c_uint64, C.c_uint]
        L.abo_extra_hash.restype = C.c_uint64
        L.abo_extra_hash.argtypes = [C.c_uint64, C.c_uint, C.c_uint]

    @staticmethod
    def _b(s):
        return s.encode() if isinstance(s, str) else bytes(s)

    def hash_seq(self, seq, k, H, mask=""):
        s = self._b(seq)
        cap = max(0, len(s) - k + 1)
        h = np.zeros((cap, H), dtype=np.uint64)
        pos = np.zeros(cap, dtype=np.uint32)
        n = self.lib.abo_hash_seq(s, len(s), k, H, mask.encode(), h.ctypes.data, pos.ctypes.data)
        return h[:n], pos[:n]

    def cbf_load(self, counters, seqs, k, H, mask=""):
        n = 0
        for s in seqs:
            s = self._b(s)
            n += self.lib.abo_cbf_load_seq(counters.ctypes.data, counters.size, s, len(s), k, H, mask.encode())
        return n

    def bf_load(self, bits, seqs, k, H, mask=""):
        n = 0
        for s in seqs:
            s = self._b(s)
            n += self.lib.abo_bf_load_seq(bits.ctypes.data, bits.size * 8, s, len(s), k, H, mask.encode())
        return n

    def casc_load(self, levels, mbits, L, seqs, k, H, mask=""):
        n = 0
        for s in seqs:
            s = self._b(s)
            n += self.lib.abo_casc_load_seq(levels.ctypes.data, mbits, L, s, len(s), k, H, mask.encode())
        return n

    def cbf_insert_hashes(self, counters, hashes):
        h = np.ascontiguousarray(hashes, dtype=np.uint64)
        H = h.shape[1]
        for i in range(h.shape[0]):
            self.lib.abo_cbf_insert(counters.ctypes.data, counters.size, h[i].ctypes.data, H)

    def cbf_min_hashes(self, counters, hashes):
        h = np.ascontiguousarray(hashes, dtype=np.uint64)
        return np.array([self.lib.abo_cbf_min(counters.ctypes.data, counters.size, h[i].ctypes.data, h.shape[1])
                         for i in range(h.shape[0])], dtype=np.uint8)


def load():
    global _lib
    if _lib is None:
        _lib = Oracle(C.CDLL(build()))
    return _lib
