"""CPU-only: the C-ABI library builds (nvcc cross-compiles), loads, and exports every symbol that
include/abyss_b200.h declares; without a GPU every compute entry point fails loudly (no fallback)."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "abyss_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(abb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(abb):
    lib = abb.load()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/abyss_b200.h but not exported"
        assert n in abb.SIGNATURES, f"{n} has no ctypes signature in abyss_b200/capi.py"
    assert lib.abb_version() == 100


def test_no_cpu_fallback(abb):
    import torch
    if torch.cuda.is_available():
        return  # the gpu suite covers the working path
    lib = abb.load()
    assert lib.abb_device_count() < 0 or lib.abb_device_count() == 0
    try:
        abb.Filter.counting(1024, 4, 20)
    except abb.AbbError as e:
        assert e.code == abb.ABB_ENODEV
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("filter creation must fail without a CUDA device")
    try:
        abb.hash_reads(5, ["ACGTACGT"])
    except abb.AbbError as e:
        assert e.code == abb.ABB_ENODEV
    else:
        raise AssertionError("hash_reads must fail without a CUDA device")
