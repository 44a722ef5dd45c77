"""Shared helpers for the test-suite (the oracle is imported ONLY from tests/)."""
import base64
import json
import os
import random

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_extract_cases():
    with open(os.path.join(GOLDEN, "extract_cases.json")) as f:
        cases = json.load(f)
    for c in cases:
        c["fasta"] = base64.b64decode(c["fasta_b64"])
    return cases


def golden_dupk_cases():
    """k_values that list a k more than once (tests/golden/make_golden.py: dupk_cases)."""
    with open(os.path.join(GOLDEN, "extract_dupk_cases.json")) as f:
        cases = json.load(f)
    for c in cases:
        c["fasta"] = base64.b64decode(c["fasta_b64"])
    return cases


def golden_metadata_cases():
    """(cases, big): genome tallies and k-mer count summaries written by the reference's metadata managers."""
    import gzip
    with open(os.path.join(GOLDEN, "metadata_cases.json")) as f:
        meta = json.load(f)
    by_name = {c["name"]: c for c in golden_extract_cases()}
    for c in meta["cases"]:
        c["fasta"] = by_name[c["name"]]["fasta"]
        c["files"] = by_name[c["name"]]["files"]
    for c in meta["big"]:
        c["fasta"] = gzip.decompress(base64.b64decode(c["fasta_gz_b64"]))
    return meta["cases"], meta["big"]


def parse_kmer_file(text):
    """k{k}.txt text -> list of (digits, count) in file order."""
    out = []
    for line in text.splitlines():
        d, c = line.split("\t")
        out.append((d, int(c)))
    return out


def fuzz_fasta(rng: random.Random, max_records=6, max_len=400):
    """Random FASTA with the corner cases the walker must get exactly right."""
    parts = []
    if rng.random() < 0.2:
        parts.append(rng.choice(["junk\n", "\n", "ACGT\n; x\n", " >notheader\n"]))
    for _ in range(rng.randint(0, max_records)):
        eol = rng.choice(["\n", "\n", "\r\n", "\r"])
        parts.append(">" + "".join(rng.choice("ACGT>x y\t") for _ in range(rng.randint(0, 150))) + eol)
        n = rng.choice([0, 1, 2, 5, 11, 12, 13, rng.randint(0, max_len), rng.randint(0, max_len)])
        alpha = rng.choice(["ACGT", "ACGT", "ACGTN", "ACGTacgtnNRY>-", "AC", "ACGT \t"])
        seq = "".join(rng.choice(alpha) for _ in range(n))
        if rng.random() < 0.3:
            i = rng.randint(0, max(0, len(seq) - 1))
            seq = seq[:i] + "N" * rng.randint(1, 40) + seq[i:]
        width = rng.choice([1, 2, 3, 7, 60, 61, 63, 64, 65, 80, 10 ** 6])
        for i in range(0, len(seq), width):
            line = seq[i:i + width]
            if rng.random() < 0.15:
                line += rng.choice([" ", "\t", " \t ", "\x0b", "\x0c", "\x1c"])
            parts.append(line + eol)
        if rng.random() < 0.2:
            parts.append(eol * rng.randint(1, 3))
    t = "".join(parts)
    if rng.random() < 0.3:
        t = t.rstrip("\r\n")
    return t.encode("latin-1")


def as_u8(data):
    return np.frombuffer(bytes(data), dtype=np.uint8)


def unwrapped_fasta_with_windows(seqs, k, canonical=False):
    """FASTA bytes with one unwrapped line per record, plus every valid k-window as (2-bit packed code uint64,
    byte offset of its last base) -- a numpy restatement of the window rules (generate.py:49-58) for this simple
    layout, used to check the byte-range / merge units of the multi-GPU sparse path."""
    parts, spans, pos = [], [], 0
    for i, s in enumerate(seqs):
        h = f">r{i}\n".encode()
        parts += [h, s, b"\n"]
        spans.append((pos + len(h), pos + len(h) + len(s)))
        pos += len(h) + len(s) + 1
    data = b"".join(parts)
    a = np.frombuffer(data, np.uint8)
    lut = np.full(256, 255, np.uint8)
    for j, ch in enumerate(b"ACGT"):
        lut[ch] = j
        lut[ch + 32] = j
    keys, ends = [], []
    for s0, s1 in spans:
        if s1 - s0 < k:
            continue
        c = lut[a[s0:s1]]
        bad = (c == 255).astype(np.int64)
        cs = np.concatenate(([0], np.cumsum(bad)))
        n = s1 - s0 - k + 1
        ok = (cs[k:k + n] - cs[:n]) == 0
        code = np.zeros(n, np.uint64)
        rc = np.zeros(n, np.uint64)
        cc = np.where(c == 255, 0, c).astype(np.uint64)
        for j in range(k):
            code = (code << np.uint64(2)) | cc[j:j + n]
            rc = rc | ((np.uint64(3) - cc[j:j + n]) << np.uint64(2 * j))
        key = np.minimum(code, rc) if canonical else code
        keys.append(key[ok])
        ends.append((np.arange(n) + s0 + k - 1)[ok])
    keys = np.concatenate(keys) if keys else np.zeros(0, np.uint64)
    ends = np.concatenate(ends) if ends else np.zeros(0, np.int64)
    return data, keys, ends.astype(np.int64)


def reduce_windows(keys, ends, begin=0, end=None):
    """(distinct keys ascending, counts, smallest end offset) of the windows whose end lies in [begin, end)."""
    sel = ends >= begin
    if end is not None:
        sel &= ends < end
    k, e = keys[sel], ends[sel]
    order = np.lexsort((e, k))
    k, e = k[order], e[order]
    uniq, idx, cnt = np.unique(k, return_index=True, return_counts=True)
    return uniq, cnt.astype(np.int64), e[idx]


def golden_big_cases():
    """Inputs of a few hundred kilobases run through the unmodified reference (tests/golden/make_golden.py:
    big_cases): FASTA bytes + SHA-256 / line count of every k{k}.txt it wrote."""
    import gzip
    with open(os.path.join(GOLDEN, "extract_big_cases.json")) as f:
        cases = json.load(f)
    for c in cases:
        c["fasta"] = gzip.decompress(base64.b64decode(c["fasta_gz_b64"]))
    return cases


def encode_reference(data):
    """Plain restatement of stage 1 for kmerml_encode: 0..3 for A C G T (either case) inside records, 255 elsewhere
    (header lines, line ends, other letters, text before the first '>' -- the SeqIO rule of SURVEY 8c).  Checked
    against the oracle's k = 1 counts in test_host_layer.py."""
    import numpy as np
    out = np.full(len(data), 255, np.uint8)
    code = {65: 0, 67: 1, 71: 2, 84: 3, 97: 0, 99: 1, 103: 2, 116: 3}
    line_start, in_header, in_record = True, False, False
    for i, ch in enumerate(data):
        if ch in (10, 13):
            line_start, in_header = True, False
            continue
        if line_start and ch == 62:
            in_header = in_record = True
        line_start = False
        if in_header or not in_record:
            continue
        c = code.get(ch)
        if c is not None:
            out[i] = c
    return out
