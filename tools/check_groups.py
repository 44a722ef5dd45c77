#!/usr/bin/env python
"""More genomes than one payload group holds (partition path, k=1..12): rows of a big batch must equal the rows of
single-genome calls, also for genomes of the later groups (tail lists / overflow lists are addressed per group)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmerml_b200 import engine
import bench

n_gen = int(sys.argv[1]) if len(sys.argv) > 1 else 130
dev = torch.device("cuda", 0)
parts, offs = [], [0]
for i in range(n_gen):
    f, nb = bench.make_genome_gpu(1000 + i, 1.0, dev, torch)
    if i % 7 == 3:                                        # some repeats and N runs: overflow + many tails
        f[1_000_000:1_400_000] = f[1_000_000:1_000_200].repeat(2000)
        seg = f[3_000_000:3_050_000]
        seg[(seg != 10) & (seg != 62) & (torch.arange(seg.numel(), device=dev) % 37 == 0)] = ord("N")
    parts.append(f); offs.append(offs[-1] + int(f.numel()))
fasta = torch.cat(parts)
ks = list(range(1, 13))
res = engine.count_dense_device(fasta, offs, ks)
torch.cuda.synchronize()
bad = 0
for g in sorted(set([0, 3, n_gen // 2, n_gen - 27, n_gen - 4, n_gen - 1])):
    one = engine.count_dense_device(parts[g], [0, int(parts[g].numel())], ks)
    same = torch.equal(one.counts[0], res.counts[g]) and torch.equal(one.totals[0], res.totals[g]) and \
        torch.equal(one.freq[0], res.freq[g])
    print(f"genome {g}: batch row == single-genome row: {same}", flush=True)
    bad += 0 if same else 1
print("payload GB (k=12):", sum(int(p.numel()) for p in parts) * 4 / 1e9, "groups of <= 12 GB")
sys.exit(1 if bad else 0)
