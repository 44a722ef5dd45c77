"""Synthetic genomes of the shapes BASELINE.json names.

Upper-case ACGT, 80-column FASTA, one header per record -- the format the
reference's downloader produces (kmerml/download/ncbidownload.py:134-153).
"""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def fasta_bytes(record_lengths, seed, width=80, probs=None, name="chr", n_runs=None):
    """-> uint8 array with the FASTA text of one genome."""
    rng = np.random.default_rng(seed)
    parts = []
    for i, L in enumerate(record_lengths):
        parts.append(np.frombuffer(f">{name}{i + 1} synthetic record len={L}\n".encode(), dtype=np.uint8))
        if probs is None:
            codes = rng.integers(0, 4, size=L, dtype=np.uint8)
        else:
            codes = rng.choice(4, size=L, p=probs).astype(np.uint8)
        seq = _ACGT[codes]
        if n_runs:
            for (start, length) in n_runs(rng, L):
                seq[start:start + length] = ord("N")
        full = (L // width) * width
        body = seq[:full].reshape(-1, width)
        lines = np.concatenate([body, np.full((body.shape[0], 1), 10, np.uint8)], axis=1).reshape(-1)
        parts.append(lines)
        if L > full:
            parts.append(seq[full:])
            parts.append(np.array([10], np.uint8))
    return np.concatenate(parts) if parts else np.zeros(0, np.uint8)


def split_lengths(total, n_records, rng):
    cuts = np.sort(rng.integers(1, total, size=n_records - 1)) if n_records > 1 else np.array([], int)
    edges = np.concatenate([[0], cuts, [total]])
    return [int(x) for x in np.diff(edges) if x > 0]


# S. cerevisiae R64 chromosome lengths (GCF_000146045.2), 16 nuclear + chrM = 12 157 105 bp
YEAST = [230218, 813184, 316620, 1531933, 576874, 270161, 1090940, 562643, 439888, 745751,
         666816, 1078177, 924431, 784333, 1091291, 948066, 85779]


def config1():
    """C1: single yeast-sized genome, k=6."""
    return fasta_bytes(YEAST, seed=1)


def config2_genome(i, scale=1.0):
    """C2: genome i of 100 fungal-sized genomes, 12-40 Mbp in 8..30 records (seed 1000+i)."""
    rng = np.random.default_rng(1000 + i)
    total = int(rng.integers(12_000_000, 40_000_000) * scale)
    nrec = int(rng.integers(8, 31))
    return fasta_bytes(split_lengths(max(total, nrec + 1), nrec, rng), seed=(1000 + i) * 7919)


def config3_genome(i, scale=1.0):
    """C3: genome i of 1000 bacterial-sized genomes, 5 Mbp, 1-3 records, GC in [0.30, 0.70]."""
    rng = np.random.default_rng(2000 + i)
    nrec = int(rng.integers(1, 4))
    gc = rng.uniform(0.30, 0.70)
    probs = [(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2]
    total = int(5_000_000 * scale)
    return fasta_bytes(split_lengths(total, nrec, rng), seed=(2000 + i) * 7919, probs=probs)


def pack(genomes):
    """Concatenate genome byte arrays back to back -> (uint8 array, n+1 byte offsets).
    No alignment is needed: the kernels clip their 64-byte chunks to each genome."""
    offs = [0]
    for g in genomes:
        offs.append(offs[-1] + len(g))
    buf = np.concatenate(genomes) if genomes else np.zeros(0, np.uint8)
    return buf, offs
