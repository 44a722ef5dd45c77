#!/usr/bin/env python
"""Development check of the tcgen05 int8 Gram kernel against numpy int64 (exact)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmerml_b200 import engine          # noqa: E402

rng = np.random.default_rng(0)
for n, m, hi in ((5, 64, 200), (130, 256, 255), (200, 4096, 70000), (300, 65536, 300), (1000, 65536, 200)):
    c = rng.integers(0, hi, size=(n, m), dtype=np.int64)
    x = torch.from_numpy(c.astype(np.uint32).view(np.int32)).cuda()
    t0 = time.perf_counter()
    d = engine.pairwise_distance_device(x, "cosine", out_dtype=torch.float64)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    g = (c @ c.T).astype(np.float64) if n * m < 3e7 else None
    if g is None:
        cf = c.astype(np.float64)
        g = cf @ cf.T
    nrm = np.sqrt(np.diag(g))
    ref = 1.0 - g / (nrm[:, None] * nrm[None, :])
    np.fill_diagonal(ref, 0.0)
    err = np.abs(d.cpu().numpy() - ref)
    off = ~np.eye(n, dtype=bool)
    rel = (err[off] / np.maximum(np.abs(ref[off]), 1e-300)).max()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    engine.pairwise_distance_device(x, "cosine", out_dtype=torch.float64)
    b.record()
    torch.cuda.synchronize()
    print(f"n={n} m={m} max count {hi}: max rel err {rel:.2e}  first call {dt * 1e3:.1f} ms, warm {a.elapsed_time(b):.2f} ms", flush=True)
