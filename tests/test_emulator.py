"""CPU emulation of the GPU thread decomposition (tests/emu/emu_dense.cpp compiles
kmerml_b200/csrc/fasta_walk.cuh with g++) against the oracle.  No GPU."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

import oracle
from helpers import fuzz_fasta, golden_extract_cases

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(HERE, "emu", "emu_dense.cpp")
    so = os.path.join(HERE, "emu", "libemu_dense.so")
    hdr = os.path.join(HERE, "..", "kmerml_b200", "csrc", "fasta_walk.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-o", so, src])
    L = ctypes.CDLL(so)
    L.emu_count_dense.restype = ctypes.c_int64
    L.emu_count_dense.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_int,
                                  ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    return L


def run_emu(L, data, kmax, min_rec, tpt, tps, base_off):
    a = np.frombuffer(data, np.uint8) if len(data) else np.zeros(0, np.uint8)
    out = np.zeros(sum(4 ** j for j in range(1, kmax + 1)), np.uint64)
    L.emu_count_dense(a.ctypes.data if a.size else None, a.size, base_off, kmax, min_rec, tpt, tps,
                      out.ctypes.data, None)
    res, off = {}, 0
    for j in range(1, kmax + 1):
        res[j] = out[off:off + 4 ** j]
        off += 4 ** j
    return res


def check(L, data, kmax, min_rec, tpt, tps, base_off):
    res = run_emu(L, data, kmax, min_rec, tpt, tps, base_off)
    for j in range(1, kmax + 1):
        ref = oracle.count_dense(data, j, min_rec)
        assert np.array_equal(ref, res[j]), (j, kmax, min_rec, tpt, tps, base_off)


def test_emulator_on_goldens(emu):
    rng = random.Random(11)
    for c in golden_extract_cases():
        for kmax in (1, 3, 8):
            check(emu, c["fasta"], kmax, kmax, rng.choice([1, 2, 4, 8]), rng.choice([1, 2, 5]),
                  rng.choice([0, 1, 17, 63, 64, 100]))


def test_emulator_fuzz(emu):
    rng = random.Random(7)
    for _ in range(400):
        data = fuzz_fasta(rng)
        kmax = rng.choice([1, 2, 3, 4, 6, 8, 9])
        mr = kmax if rng.random() < 0.75 else kmax + rng.randint(1, 12)
        check(emu, data, kmax, mr, rng.choice([1, 2, 3, 4, 8, 16]), rng.choice([1, 2, 3, 100]),
              rng.choice([0, 0, 3, 31, 64, 77]))


def test_header_text_that_looks_like_sequence(emu):
    """A header line made of ACGT letters whose line feed closes the 32-byte chunk right before a
    tile / slice start must not leak into the next record's first windows."""
    for tpt in (1, 2, 4):
        for base_off in (0, 32, 64):
            for hdrlen in range(20, 110, 9):
                hdr = (">" + "ACGT" * 40)[:hdrlen]
                data = (">r0\n" + "ACGTTGCA" * 12 + "\n" + hdr + "\n" + "GATTACAGATTACAGGATCC" * 8 + "\n").encode()
                for k in (12,):
                    check(emu, data, k, k, tpt, 1, base_off)


def test_counting_one_level_above_the_largest_k(emu):
    """k = 8 is counted as 9-mers (min_rec stays 8) and cascaded down: every level <= 8 must still be
    exactly the reference's count with max(k_values) = 8."""
    rng = random.Random(21)
    for _ in range(120):
        data = fuzz_fasta(rng)
        res = run_emu(emu, data, 9, 8, rng.choice([1, 2, 4, 16]), rng.choice([1, 3]), rng.choice([0, 31, 64]))
        for j in range(1, 9):
            assert np.array_equal(oracle.count_dense(data, j, 8), res[j]), j
        assert np.array_equal(oracle.count_dense(data, 9, 9), res[9])


def test_long_lines_and_bounded_lookback(emu):
    """Unwrapped FASTA: slice starts whose bounded look-back finds no line terminator are settled by the
    slice-table passes.  A tiny limit forces that route on ordinary inputs, too."""
    emu.emu_set_line_limit.argtypes = [ctypes.c_uint64]
    rng = random.Random(33)
    try:
        for limit in (32, 40, 64):
            emu.emu_set_line_limit(limit)
            for _ in range(120):
                data = fuzz_fasta(rng)
                kmax = rng.choice([1, 3, 6, 9])
                check(emu, data, kmax, kmax, rng.choice([1, 2, 3, 4]), rng.choice([1, 2, 3]), rng.choice([0, 3, 31, 64]))
            for c in golden_extract_cases():
                check(emu, c["fasta"], 8, 8, rng.choice([1, 2, 4]), rng.choice([1, 2]), rng.choice([0, 17, 64]))
            # long sequence lines, long header lines (some made of base letters), CRLF
            for _ in range(40):
                parts = []
                for r in range(rng.randint(1, 4)):
                    hl = rng.choice([3, 30, 200, 700])
                    alphabet = "ACGT" if rng.random() < 0.5 else "ACGTxyz _|"
                    parts.append(">" + "".join(rng.choice(alphabet) for _ in range(hl)))
                    for _line in range(rng.randint(0, 3)):
                        parts.append("".join(rng.choice("ACGTACGTACGTN") for _ in range(rng.choice([0, 5, 60, 400, 1500]))))
                nl = rng.choice(["\n", "\r\n"])
                data = (nl.join(parts) + (nl if rng.random() < 0.8 else "")).encode()
                kmax = rng.choice([2, 7, 9])
                check(emu, data, kmax, kmax, rng.choice([1, 2, 4]), rng.choice([1, 2, 3]), rng.choice([0, 5, 32, 64]))
    finally:
        emu.emu_set_line_limit(0)


def test_slices_that_are_not_whole_tiles(emu):
    """count8_kernel walks 32 KB tiles over slices made of 16 KB units: the last tile of a slice is clipped to the
    slice end.  Emulated with small numbers: tiles of 4 or 8 chunks, slices of 3, 5, 7 ... chunks."""
    emu.emu_set_slice_bytes.argtypes = [ctypes.c_uint64]
    rng = random.Random(44)
    try:
        for _ in range(250):
            tpt = rng.choice([4, 8])
            emu.emu_set_slice_bytes(32 * rng.choice([1, 3, 5, 6, 7, 9, 11, 13]))
            data = fuzz_fasta(rng)
            kmax = rng.choice([1, 3, 6, 8])
            check(emu, data, kmax, kmax, tpt, 1, rng.choice([0, 3, 32, 64]))
        for c in golden_extract_cases():
            emu.emu_set_slice_bytes(32 * rng.choice([3, 5, 7]))
            check(emu, c["fasta"], 8, 8, 4, 1, rng.choice([0, 17]))
    finally:
        emu.emu_set_slice_bytes(0)


def test_emulator_on_big_reference_cases(emu):
    """The GPU thread decomposition (real tile size: 512 chunks, 8 tiles per slice) on the big reference-pinned
    inputs, against the oracle (which test_oracle_golden pins to the reference's own output)."""
    from helpers import golden_big_cases
    for c in golden_big_cases():
        kmax = min(max(c["k_values"]), 9)
        check(emu, c["fasta"], kmax, kmax, 512, 8, 0)
