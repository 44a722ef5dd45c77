"""ctypes front-end of oracle/kmer_oracle.c plus numpy/pure-Python restatements
of the reference's Python-only stages.  TEST INFRASTRUCTURE (see __init__).

Citations are relative to /root/reference.
"""
import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FILE_DIGIT = {"A": "0", "T": "1", "C": "2", "G": "3"}      # generate.py:71
LEX = "ACGT"                                                # repo-wide index order


def build(force=False):
    so = os.path.join(_HERE, "libkmer_oracle.so")
    src = os.path.join(_HERE, "kmer_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libkmer_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        u8p = ctypes.c_void_p
        L.kmo_parse.restype = ctypes.c_int64
        L.kmo_parse.argtypes = [u8p, ctypes.c_uint64, u8p, u8p, u8p, u8p, ctypes.c_uint64]
        L.kmo_count_dense.restype = ctypes.c_int64
        L.kmo_count_dense.argtypes = [u8p, ctypes.c_uint64, ctypes.c_int, ctypes.c_int, u8p, u8p]
        L.kmo_count_dense_multi.restype = ctypes.c_int64
        L.kmo_count_dense_multi.argtypes = [u8p, ctypes.c_uint64, u8p, ctypes.c_int, ctypes.c_int, u8p]
        L.kmo_count_sparse.restype = ctypes.c_int64
        L.kmo_count_sparse.argtypes = [u8p, ctypes.c_uint64, ctypes.c_int, ctypes.c_int, u8p, u8p,
                                       ctypes.c_uint64, u8p]
        L.kmo_genome_stats.restype = ctypes.c_int
        L.kmo_genome_stats.argtypes = [u8p, ctypes.c_uint64, u8p, u8p, u8p, u8p]
        _LIB = L
    return _LIB


def _buf(data):
    a = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, (a.ctypes.data if a.size else None)


def parse(data):
    """-> list of (header_offset, sequence_bytes) as Bio.SeqIO.parse would yield (generate.py:39)."""
    a, p = _buf(data)
    n = lib().kmo_parse(p, a.size, None, None, None, None, 0)
    seq = np.zeros(max(a.size, 1), np.uint8)
    off = np.zeros(max(n, 1), np.uint64)
    ln = np.zeros(max(n, 1), np.uint64)
    hdr = np.zeros(max(n, 1), np.uint64)
    lib().kmo_parse(p, a.size, seq.ctypes.data, off.ctypes.data, ln.ctypes.data, hdr.ctypes.data, n)
    return [(int(hdr[i]), seq[int(off[i]):int(off[i] + ln[i])].tobytes()) for i in range(n)]


def record_ids(data):
    """record.id (first word of the title line) and length per record."""
    a, _ = _buf(data)
    raw = a.tobytes()
    out = []
    for hdr, seq in parse(a):
        end = hdr
        while end < len(raw) and raw[end] not in (10, 13):
            end += 1
        title = raw[hdr + 1:end].decode("latin-1").rstrip()
        words = title.split(None, 1)
        out.append((words[0] if words else "", len(seq)))
    return out


def count_dense(data, k, min_len=None, want_order=False):
    """Forward-strand dense counts (uint64[4^k], lex ACGT index) of generate.py:36-58.

    With want_order also returns the bins in dict-insertion order (generate.py:88)."""
    a, p = _buf(data)
    counts = np.zeros(4 ** k, np.uint64)
    first = np.zeros(4 ** k, np.uint64) if want_order else None
    w = lib().kmo_count_dense(p, a.size, k, k if min_len is None else min_len, counts.ctypes.data,
                              first.ctypes.data if want_order else None)
    if w < 0:
        raise RuntimeError("kmo_count_dense failed")
    if want_order:
        nz = np.nonzero(counts)[0]
        order = nz[np.argsort(first[nz], kind="stable")]
        return counts, order
    return counts


def count_dense_multi(data, ks, min_len=None):
    """{k: uint64[4^k]} for a list of k with the reference's max(k) record filter."""
    ks = list(ks)
    ml = max(ks) if min_len is None else min_len
    a, p = _buf(data)
    karr = np.asarray(ks, np.int32)
    row = np.zeros(sum(4 ** k for k in ks), np.uint64)
    w = lib().kmo_count_dense_multi(p, a.size, karr.ctypes.data, len(ks), ml, row.ctypes.data)
    if w < 0:
        raise RuntimeError("kmo_count_dense_multi failed")
    out, off = {}, 0
    for k in ks:
        out[k] = row[off:off + 4 ** k]
        off += 4 ** k
    return out


def count_sparse(data, k, min_len=None):
    """(codes uint64[d], counts uint64[d]) in dict-insertion order, any k <= 32."""
    a, p = _buf(data)
    cap = max(a.size, 1)
    codes = np.zeros(cap, np.uint64)
    counts = np.zeros(cap, np.uint64)
    win = ctypes.c_uint64(0)
    d = lib().kmo_count_sparse(p, a.size, k, k if min_len is None else min_len, codes.ctypes.data,
                               counts.ctypes.data, cap, ctypes.addressof(win))
    if d < 0:
        raise RuntimeError("kmo_count_sparse failed")
    return codes[:d].copy(), counts[:d].copy()


def genome_stats(data):
    a, p = _buf(data)
    v = [ctypes.c_uint64(0) for _ in range(4)]
    lib().kmo_genome_stats(p, a.size, *[ctypes.addressof(x) for x in v])
    contigs, total, gc, nn = (int(x.value) for x in v)
    return {"contigs": contigs, "total_size": total, "n_count": nn,
            "gc_content": (gc / total) * 100 if total else 0}


# ---------------------------------------------------------------- helpers
def code_to_kmer(code, k):
    return "".join(LEX[(int(code) >> (2 * (k - 1 - i))) & 3] for i in range(k))


def kmer_to_code(s):
    c = 0
    for ch in s:
        c = (c << 2) | LEX.index(ch)
    return c


def file_digits(code, k):
    """digit string written by generate.py:87-91 (A0 T1 C2 G3)."""
    return "".join(FILE_DIGIT[ch] for ch in code_to_kmer(code, k))


def kmer_file_text(data, k, min_len, multiplicity=1):
    """Exact text of k{k}.txt as generate.py:68-91 writes it.  `multiplicity`: how often k appears in
    k_values -- `for k in k_values` (generate.py:49) counts into ONE dict per distinct k (:36) once per entry."""
    if k <= 12:
        counts, order = count_dense(data, k, min_len, want_order=True)
        return "".join(f"{file_digits(b, k)}\t{int(counts[b]) * multiplicity}\n" for b in order)
    codes, counts = count_sparse(data, k, min_len)
    return "".join(f"{file_digits(c, k)}\t{int(n) * multiplicity}\n" for c, n in zip(codes, counts))


def kmer_file_stats(text, k):
    """kmerml/utils/kmer_metadata.py:59-78 for the text of one k{k}.txt (numpy restatement of the pandas
    reductions: mean = sum / n, median = middle element or the mean of the two middle ones)."""
    counts = np.array([int(line.split("\t")[1]) for line in text.splitlines()], dtype=np.int64)
    total = int(counts.sum())
    return {"k_value": k, "total_kmers": total, "unique_kmers": int(counts.size), "max_count": int(counts.max()),
            "min_count": int(counts.min()), "mean_count": float(counts.mean()), "median_count": float(np.median(counts)),
            "estimated_genome_size": total + k - 1}


def revcomp_code(code, k):
    rc = 0
    c = int(code)
    for _ in range(k):
        rc = (rc << 2) | (3 - (c & 3))
        c >>= 2
    return rc


def canonical_from_forward(fwd, k):
    """Canonical counts pinned THROUGH the reference's forward counts:
    canon[c] = fwd[c] + fwd[rc(c)] for c < rc(c), fwd[c] for palindromes, 0 elsewhere."""
    idx = np.arange(4 ** k, dtype=np.uint64)
    rc = np.zeros_like(idx)
    t = idx.copy()
    for _ in range(k):
        rc = (rc << np.uint64(2)) | (np.uint64(3) - (t & np.uint64(3)))
        t >>= np.uint64(2)
    out = np.zeros_like(fwd)
    lo = idx < rc
    out[lo] = fwd[lo] + fwd[rc[lo]]
    pal = idx == rc
    out[pal] = fwd[pal]
    return out


def frequencies(counts):
    """float64 row / row-sum (the normalisation the reference only gestures at, tests/test_ml.py:8)."""
    c = np.asarray(counts, np.float64)
    s = c.sum()
    return c / s if s > 0 else c


def pairwise_distance(X, metric):
    """float64 genome x genome distances on rows of X (oracle for SURVEY 8a row 13)."""
    X = np.asarray(X, np.float64)
    if metric == "euclidean":
        d2 = ((X[:, None, :] - X[None, :, :]) ** 2).sum(-1) if X.shape[0] * X.shape[0] * X.shape[1] < 5e7 else None
        if d2 is None:
            g = X @ X.T
            n = np.diag(g)
            d2 = np.maximum(n[:, None] + n[None, :] - 2 * g, 0)
        return np.sqrt(d2)
    if metric == "cosine":
        g = X @ X.T
        n = np.sqrt(np.diag(g))
        with np.errstate(divide="ignore", invalid="ignore"):
            d = 1.0 - g / (n[:, None] * n[None, :])
        np.fill_diagonal(d, 0.0)
        return d
    raise ValueError(metric)


# ------------------------------------------------- statistics.py restatement
def decode_digits(encoded):
    """statistics.py:248-251 applied to str(int) -- leading zeros (= leading A) are already lost."""
    m = {"0": "A", "1": "T", "2": "C", "3": "G"}
    return "".join(m.get(c, "N") for c in str(encoded))


def compat_kmer_string(code, k):
    """What the reference's CSV ends up holding for a k-mer: digits -> int -> str -> decode
    (statistics.py:158, 261-271 int inference)."""
    return decode_digits(int(file_digits(code, k)))


def kmer_features(s, required=None):
    """statistics.py:149-240 for one (possibly truncated) k-mer string; returns ordered dict items."""
    if required is None:
        required = ["base_counts", "gc_content", "cpg_sites", "entropy", "repeats", "presence"]
    f = {}
    n = len(s)
    if "gc_content" in required:
        gc = s.count("G") + s.count("C")
        f["gc_percent"] = (gc / n) * 100 if n > 0 else 0
    if "base_counts" in required:
        for b in "ACGT":
            f[f"{b}_count"] = s.count(b)
    if "presence" in required:
        for b in "ACGT":
            f[f"{b}_present"] = 1 if b in s else 0
    if "cpg_sites" in required:
        cpg = sum(1 for i in range(n - 1) if s[i:i + 2] == "CG")
        f["cpg_count"] = cpg
        cf = s.count("C") / n if n > 0 else 0
        gf = s.count("G") / n if n > 0 else 0
        expected = cf * gf * (n - 1) if cf * gf > 0 else 0.001
        f["cpg_obs_exp"] = cpg / expected if expected > 0 else 0
    if "entropy" in required:
        ent = 0
        for b in set(s):
            p = s.count(b) / n
            ent -= p * math.log2(p) if p > 0 else 0
        f["shannon_entropy"] = ent
        f["normalized_entropy"] = ent / 2.0
    if "repeats" in required:
        f["has_repeat"] = 0
        for i in range(n - 3):
            if s[i:i + 2] == s[i + 2:i + 4]:
                f["has_repeat"] = 1
                break
    return f
