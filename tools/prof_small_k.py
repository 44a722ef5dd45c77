#!/usr/bin/env python
"""C3-shaped counting for the ncu captures of the k <= 8 kernels: N genomes x 5 Mbp, k = 8 then k = 7 (and 6).
    python tools/prof_small_k.py [n_genomes]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench_extras
from kmerml_b200 import engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
dev = torch.device("cuda", 0)
parts, offs = [], [0]
for i in range(n):
    g = bench_extras.gpu_fasta(torch, dev, [5_000_000], 2000 + i, f"b{i}", gc=0.3 + 0.4 * (i % 7) / 6)
    parts.append(g); offs.append(offs[-1] + g.numel())
fasta = torch.cat(parts); del parts
for ks in ([8], [7], [6]):
    for rep in range(3):
        torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); res = engine.count_dense_device(fasta, offs, ks, want_freq=True); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"k={ks} {n} x 5 Mbp: {ms:.3f} ms  {n * 5e6 / ms / 1e6:.1f} Gbp/s", flush=True)
