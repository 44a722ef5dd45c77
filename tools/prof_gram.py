#!/usr/bin/env python
"""The C3-sized Gram / distance computation for the ncu capture: 1000 genomes x 65536 features (k = 8 count rows of
5 Mbp genomes: counts around 76, two digit planes), kmerml_pairwise_distance on the tensor cores."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from kmerml_b200 import engine

rng = np.random.default_rng(0)
x = torch.from_numpy(rng.poisson(76, size=(1000, 65536)).astype(np.int32)).cuda()
for _ in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); d = engine.pairwise_distance_device(x, "cosine"); b.record(); torch.cuda.synchronize()
print(f"1000 x 65536: {a.elapsed_time(b):.3f} ms", flush=True)
