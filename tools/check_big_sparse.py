#!/usr/bin/env python
"""BASELINE config 5 in spirit on ONE GPU: a synthetic genome of KM_GBP Gbp (default 1.0) with N runs,
k = 21 canonical through the sparse path (emit -> radix sort -> run-length reduce).  Size-independent checks:
sum(counts) = windows, k-mers strictly ascending, every k-mer canonical, first offsets inside the file."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmerml_b200 import engine, synth              # noqa: E402
import bench                                       # noqa: E402


def main():
    gbp = float(os.environ.get("KM_GBP", "1.0"))
    k = int(os.environ.get("KM_K", "21"))
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(5)
    lens = synth.split_lengths(int(gbp * 1e9), 24, rng)
    bench.genome_shape = lambda i, scale: lens
    fasta, nbases = bench.make_genome_gpu(5, 1.0, dev, torch)
    # N runs: 200 random runs of 100..50000 bases (line feeds and header bytes are kept)
    n = int(fasta.numel())
    for a, ln in zip(rng.integers(1000, n - 60_000, 200).tolist(), rng.integers(100, 50_000, 200).tolist()):
        seg = fasta[a:a + ln]
        seg[(seg == 65) | (seg == 67) | (seg == 71) | (seg == 84)] = ord("N")
    print(f"genome: {nbases / 1e9:.3f} Gbp, {n / 1e9:.3f} GB FASTA, k={k} canonical", flush=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    keys, counts, first, windows = engine.count_sparse_device(fasta, k, canonical=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"sparse count: {dt * 1e3:.0f} ms ({nbases / dt / 1e9:.2f} Gbp/s incl. allocation), {keys.numel()} distinct, "
          f"{windows} windows", flush=True)
    total, step, prev_last = 0, 100_000_000, None
    for a in range(0, keys.numel(), step):                      # chunked: the outputs alone are tens of GB
        kc = keys[a:a + step]
        total += int(counts[a:a + step].to(torch.int64).sum().item())
        assert bool((kc[1:] > kc[:-1]).all().item()), "k-mers not strictly ascending"
        if prev_last is not None:
            assert int(kc[0].item()) > prev_last
        prev_last = int(kc[-1].item())
        f = first[a:a + step].to(torch.int64) & 0xFFFFFFFF
        assert int(f.max().item()) < n and int(f.min().item()) >= k - 1
    assert total == windows, (total, windows)
    # canonical: key <= reverse complement (sample)
    idx = torch.randint(0, keys.numel(), (1_000_000,), device=dev)
    s = keys[idx]
    rc = torch.zeros_like(s)
    t = s.clone()
    for _ in range(k):
        rc = (rc << 2) | (3 - (t & 3))
        t = t >> 2
    assert bool((s <= rc).all().item()), "non-canonical k-mer in the output"
    del keys, counts, first
    torch.cuda.empty_cache()
    t0 = time.perf_counter()
    out = engine.count_sparse_device(fasta, k, canonical=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"second call: {dt * 1e3:.0f} ms  {nbases / dt / 1e9:.2f} Gbp/s; checks passed", flush=True)


if __name__ == "__main__":
    main()
