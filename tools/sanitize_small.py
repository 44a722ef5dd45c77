#!/usr/bin/env python
"""A small tour through every kernel family on tiny inputs, each result compared with the oracle.  Small enough to run
under compute-sanitizer where that is available (it is closed on the round-1 GPU pool, so round 1 ran it plain)."""
import os, random, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
from helpers import fuzz_fasta, golden_extract_cases
from kmerml_b200 import engine

def dev_of(b):
    a = np.frombuffer(b, np.uint8) if len(b) else np.zeros(0, np.uint8)
    return torch.from_numpy(np.concatenate([a, np.zeros(64, np.uint8)])).cuda()[:a.size]

rng = random.Random(5)
cases = [c["fasta"] for c in golden_extract_cases()][:12] + [fuzz_fasta(rng) for _ in range(12)]
seq = "".join(rng.choice("ACGT") for _ in range(90_000))
cases.append((">long one line\n" + seq + "\n>r2\n" + "ACGTTGCAAT" * 3000 + "N" + seq[:700] + "\n").encode())
cases.append((">busy\n" + "N".join(seq[i:i + 11] for i in range(0, 40_000, 11)) + "\n").encode())
n_ok = 0
for data in cases:
    if not len(data):
        continue
    d = dev_of(data)
    for ks in ([1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12], [8], [6], [10], [13]):
        res = engine.count_dense_device(d, [0, d.numel()], ks)
        for k in (ks[-1], ks[0]):
            ref = oracle.count_dense(data, k, max(ks))
            assert np.array_equal(ref, res.counts_numpy(0, k).astype(np.uint64)), (ks, k)
        n_ok += 1
    res = engine.count_dense_device(d, [0, d.numel()], [9, 12], canonical=True)
    first = engine.first_occurrence_device(d, 9)
    engine.format_kmer_file_device(engine.count_dense_device(d, [0, d.numel()], [9], want_freq=False).counts_of(0, 9), first, 9)
    keys, cnts, fst, _ = engine.count_sparse_device(d, 17)
    wk, wc = oracle.count_sparse(data, 17)
    assert np.array_equal(np.sort(wk), keys.cpu().numpy().view(np.uint64))
    if keys.numel():
        engine.format_kmer_lines_device(keys, cnts, 17)
        engine.merge_sparse_device(torch.cat([keys, keys]), torch.cat([cnts, cnts]), torch.cat([fst, fst]), 17)
    engine.genome_stats_device(d)
x = engine.count_dense_device(dev_of(cases[-2]), [0, len(cases[-2])], [8]).counts
engine.pairwise_distance_device(torch.cat([x, x, x]), "cosine")
engine.static_features_device(6)
torch.cuda.synchronize()
print("sanitize tour ok:", n_ok, "dense calls")
