#!/usr/bin/env python
"""Device-resident timings of the other BASELINE configs (development tool, not the bench contract):
C1 = one yeast-sized genome k=6; C3 = n bacterial-sized genomes k=8 (+ cosine distance matrix)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmerml_b200 import engine, synth          # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


def main():
    n3 = int(os.environ.get("KM_C3", "200"))
    g1 = torch.from_numpy(synth.config1()).cuda()
    nb1 = 12_157_105
    for ks in ([6], [8], [12], list(range(1, 13))):
        ms = timed(lambda: engine.count_dense_device(g1, [0, g1.numel()], ks))
        print(f"C1 12.16 Mbp k={ks if len(ks) < 4 else '1..12'}: {ms:.3f} ms  {nb1 / ms / 1e6:.1f} Gbp/s")
    gs = [synth.config3_genome(i) for i in range(n3)]
    buf, offs = synth.pack(gs)
    dev = torch.from_numpy(buf).cuda()
    nb3 = n3 * 5_000_000
    for ks, part in (([8], True), ([8], "k8as9"), ([7], True), ([9], True)):
        ms = timed(lambda: engine.count_dense_device(dev, offs, ks, partition=part), reps=3)
        print(f"C3 {n3} x 5 Mbp k={ks} ({part}): {ms:.2f} ms  {nb3 / ms / 1e6:.1f} Gbp/s")
    res = engine.count_dense_device(dev, offs, [8])
    ms = timed(lambda: engine.pairwise_distance_device(res.counts, "cosine"), reps=3)
    print(f"C3 cosine distance {n3} x {n3} over 65536 features (fp64 Gram): {ms:.2f} ms")


if __name__ == "__main__":
    main()
