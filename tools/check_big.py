#!/usr/bin/env python
"""BASELINE config 4 in spirit on ONE GPU: a single synthetic genome of KM_GBP Gbp (default 1.0), 24 records,
k = 12 forward and canonical.  Full-size parity through size-independent properties: sum(counts) = windows,
k=1 total = bases, byte ranges add up to the whole, canonical = fold of forward, and a bit-exact oracle
comparison on a 20 Mbp genome cut from the same stream."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle                                      # noqa: E402  (checker only)
from kmerml_b200 import dist as kdist             # noqa: E402
from kmerml_b200 import engine, synth              # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)) + "/..")
from bench import make_genome_gpu                  # noqa: E402


def main():
    gbp = float(os.environ.get("KM_GBP", "1.0"))
    dev = torch.device("cuda", 0)
    import bench
    rng = np.random.default_rng(4)
    lens = synth.split_lengths(int(gbp * 1e9), 24, rng)
    bench.genome_shape = lambda i, scale: lens           # reuse the GPU generator with our record lengths
    fasta, nbases = make_genome_gpu(4, 1.0, dev, torch)
    print(f"genome: {nbases / 1e9:.3f} Gbp, {fasta.numel() / 1e9:.3f} GB FASTA")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = engine.count_dense_device(fasta, [0, int(fasta.numel())], [12, 1])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    c12 = res.counts_of(0, 12).to(torch.int64) & 0xFFFFFFFF
    c1 = res.counts_of(0, 1).to(torch.int64) & 0xFFFFFFFF
    windows = int(c12.sum().item())
    print(f"k=12 + k=1 in {dt * 1e3:.1f} ms ({nbases / dt / 1e9:.1f} Gbp/s incl. first-call allocation)")
    assert windows == int(res.totals[0, 0].item()) == nbases - 11 * len(lens), (windows, nbases)
    assert int(c1.sum().item()) == nbases == int(res.totals[0, 1].item())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    engine.count_dense_device(fasta, [0, int(fasta.numel())], [12], want_freq=False)
    b.record()
    torch.cuda.synchronize()
    print(f"warm k=12 counts only: {a.elapsed_time(b):.1f} ms  {nbases / a.elapsed_time(b) / 1e6:.1f} Gbp/s")
    # byte ranges add up (the multi-GPU invariant), 4 ranges
    acc = torch.zeros_like(c12)
    for rb, re_ in kdist.chunk_ranges(int(fasta.numel()), 4):
        c, _ = engine.count_dense_range_device(fasta, rb, re_, [12])
        acc += c.to(torch.int64) & 0xFFFFFFFF
    assert torch.equal(acc, c12), "ranges do not add up"
    # canonical = fold of forward
    can = engine.count_dense_device(fasta, [0, int(fasta.numel())], [12], canonical=True, want_freq=False)
    cc = can.counts_of(0, 12).to(torch.int64) & 0xFFFFFFFF
    rc = torch.as_tensor(engine.revcomp_codes(12), device=dev)
    idx = torch.arange(4 ** 12, device=dev)
    want = torch.where(idx < rc, c12 + c12[rc], torch.where(idx == rc, c12, torch.zeros_like(c12)))
    assert torch.equal(cc, want), "canonical fold mismatch"
    # bit-exact oracle comparison on a 20 Mbp piece
    cut = 0
    raw = fasta[:25_000_000].cpu().numpy()
    nl = np.nonzero(raw == 10)[0]
    cut = int(nl[nl < 20_300_000][-1]) + 1
    piece = raw[:cut]
    got = engine.count_dense_device(fasta[:cut].contiguous(), [0, cut], [12], want_freq=False).counts_numpy(0, 12)
    ref = oracle.count_dense(piece.tobytes(), 12)
    assert np.array_equal(got.astype(np.uint64), ref), "oracle mismatch on the 20 Mbp piece"
    print("big-genome checks passed: totals, range additivity, canonical fold, oracle piece")


if __name__ == "__main__":
    main()
