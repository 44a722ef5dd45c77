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
