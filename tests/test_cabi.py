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
