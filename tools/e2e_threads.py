#!/usr/bin/env python
"""How the host-buffer call scales with the number of widening threads (and the plain uint32 copy beside it).
    python tools/e2e_threads.py [n_genomes]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from kmerml_b200 import _lib, engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev = torch.device("cuda", 0)
ks = list(range(1, 13))
hosts, nb = [], 0
for i in range(n):
    t, b = bench.make_genome_gpu(i, 1.0, dev, torch)
    h = torch.empty(t.numel(), dtype=torch.uint8, pin_memory=True)
    h.copy_(t)
    hosts.append(h)
    nb += b
_, row_len = engine.row_layout(ks)
hc = torch.empty((n, row_len), dtype=torch.int32, pin_memory=True)
ht = torch.zeros((n, len(ks)), dtype=torch.int64, pin_memory=True)
freq = torch.empty((n, row_len), dtype=torch.float32, device=dev)
ctx = _lib.context(0)
ref = None
for label, thr, wide in [("wide", 1, True)] + [(f"narrow{t}", t, False) for t in (2, 4, 8, 12, 15, 16)]:
    ctx.set_host_threads(thr)
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        engine.count_dense_host(hosts, ks, device=dev, out_counts=hc, out_freq=freq, out_totals=ht, wide_d2h=wide)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    if ref is None:
        ref = hc.clone()
    print(f"{label:9s} {dt * 1e3:8.1f} ms  {nb / dt / 1e9:6.1f} Gbp/s  host rows written {n * row_len * 4 / dt / 1e9:6.1f} GB/s  same={torch.equal(ref, hc)}", flush=True)
