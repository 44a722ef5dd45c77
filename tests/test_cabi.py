"""The C-ABI library loads and exports every symbol include/kmerml_b200.h declares;
without a GPU every compute entry point fails loudly (no CPU fallback).  No GPU."""
import ctypes
import os
import re

import pytest

from kmerml_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "kmerml_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kmerml_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 8
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/kmerml_b200.h but not exported"
    assert sorted(_lib.EXPORTS) == names


def test_version_and_row_len(lib):
    assert lib.kmerml_version() >= 100
    ks = (ctypes.c_int * 3)(1, 2, 12)
    assert lib.kmerml_row_len(ks, 3) == 4 + 16 + 4 ** 12


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = ctypes.c_void_p()
    rc = lib.kmerml_ctx_create(0, ctypes.byref(h))
    assert rc != 0 and not h
    assert b"no CUDA device" in lib.kmerml_last_error()
    with pytest.raises(_lib.KmermlError):
        _lib.Context(0)
    from kmerml_b200 import engine
    with pytest.raises(_lib.KmermlError):
        engine.count_dense_host([b">a\nACGT\n"], [2])


def test_product_never_imports_oracle():
    """The product package must not reference the oracle (test infrastructure)."""
    pkg = os.path.join(ROOT, "kmerml_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "kmer_oracle" not in src, f


def _wire_row(ks, rows, nibble_mask, exc_cap=32768):
    """A compact count row built by hand from the format include/kmerml_b200.h documents: header ("KMW2", nibble mask) |
    byte / nibble block of the levels k >= 10 | exception count | exception list | uint32 small levels."""
    import numpy as np
    narrow, small, exc, off = [], [], [], 0
    for i, k in enumerate(ks):
        c = rows[i].astype(np.uint64)
        if k >= 10:
            if (nibble_mask >> i) & 1:
                sat = np.minimum(c, 15).astype(np.uint8)
                narrow.append((sat[0::2] | (sat[1::2] << 4)).astype(np.uint8))
                big = np.nonzero(c >= 15)[0]
            else:
                narrow.append(np.minimum(c, 255).astype(np.uint8))
                big = np.nonzero(c >= 255)[0]
            exc += [(off + int(b), int(c[b])) for b in big]
        else:
            small.append(c.astype(np.uint32))
        off += 4 ** k
    head = np.array([0x32574D4B, nibble_mask, 0, 0], np.uint32).tobytes()
    e = np.zeros((exc_cap, 2), np.uint32)
    for j, (b, v) in enumerate(exc[:exc_cap]):
        e[j] = (b, v)
    body = head + b"".join(a.tobytes() for a in narrow) + np.array([len(exc), 0, 0, 0], np.uint32).tobytes() + e.tobytes() + \
        b"".join(a.tobytes() for a in small)
    return body + b"\0" * (-len(body) % 16)


def test_compact_row_decoder_on_hand_built_rows():
    """kmerml_compact_expand / _row_bytes / _row_used_bytes / _row_overflowed are host code: check them without a GPU
    on rows assembled here from the documented wire format (byte and nibble levels, saturated bins in the exception
    list, small levels, an overflowed list)."""
    import ctypes
    import numpy as np
    from kmerml_b200 import _lib
    L = _lib.load()
    rng = np.random.default_rng(3)
    for ks, mask in (([10, 3], 0), ([10, 3], 1), ([3, 11, 10, 1], 0b0110), ([11, 10], 0b01)):
        rows = []
        for k in ks:
            c = rng.poisson(3.0, 4 ** k).astype(np.uint64)
            hot = rng.integers(0, 4 ** k, 40)
            c[hot] = rng.integers(15, 5000, 40)                       # above both saturation values
            c[int(hot[0])] = 4_000_000_000                            # needs all 32 bits
            rows.append(c)
        wire = _wire_row(ks, rows, mask)
        karr = np.asarray(ks, dtype=np.int32)
        kp, n = karr.ctypes.data, len(ks)
        buf = np.frombuffer(wire, np.uint8).copy()
        assert int(L.kmerml_compact_row_used_bytes(kp, n, buf.ctypes.data)) == len(wire)
        assert int(L.kmerml_compact_row_bytes(kp, n)) >= len(wire)     # the all-bytes layout is the largest
        assert L.kmerml_compact_row_overflowed(kp, n, buf.ctypes.data) == 0
        for i, k in enumerate(ks):
            out = np.zeros(4 ** k, np.uint32)
            assert L.kmerml_compact_expand(kp, n, buf.ctypes.data, i, out.ctypes.data) == 0
            assert np.array_equal(out.astype(np.uint64), rows[i]), (ks, mask, k)
        assert L.kmerml_compact_expand(kp, n, buf.ctypes.data, n, None) < 0          # bad level index / null output
    # an exception list that ran over: flagged, and expanding a narrow level refuses instead of returning saturated bins
    ks = [10, 2]
    rows = [np.full(4 ** 10, 300, np.uint64), np.arange(16, dtype=np.uint64)]
    wire = _wire_row(ks, rows, 0)
    karr = np.asarray(ks, dtype=np.int32)
    buf = np.frombuffer(wire, np.uint8).copy()
    assert L.kmerml_compact_row_overflowed(karr.ctypes.data, 2, buf.ctypes.data) == 1
    out = np.zeros(4 ** 10, np.uint32)
    assert L.kmerml_compact_expand(karr.ctypes.data, 2, buf.ctypes.data, 0, out.ctypes.data) < 0
    small = np.zeros(16, np.uint32)
    assert L.kmerml_compact_expand(karr.ctypes.data, 2, buf.ctypes.data, 1, small.ctypes.data) == 0
    assert np.array_equal(small, np.arange(16, dtype=np.uint32))
    bad = buf.copy()
    bad[0] ^= 0xFF
    assert L.kmerml_compact_expand(karr.ctypes.data, 2, bad.ctypes.data, 1, small.ctypes.data) < 0   # not a wire row
