#!/usr/bin/env python
"""Wall time of the drop-in KmerExtractor on one synthetic genome (development tool)."""
import os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmerml_b200 import synth
from kmerml_b200.kmers.generate import KmerExtractor

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12_000_000
ks = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else list(range(1, 13))
with tempfile.TemporaryDirectory() as tmp:
    fa = os.path.join(tmp, "GCA_000001_synthetic.fna")
    rng = np.random.default_rng(3)
    seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)]
    rows = [seq[i:i + 80].tobytes() for i in range(0, n, 80)]
    with open(fa, "wb") as f:
        f.write(b">chr1 synthetic\n" + b"\n".join(rows) + b"\n")
    for compress in (False,):
        ex = KmerExtractor(output_dir=os.path.join(tmp, f"out{int(compress)}"), compress=compress)
        ex.extract_kmers_from_fasta(fa, [6])           # warm-up (context, workspaces)
        t = time.perf_counter()
        ex.extract_kmers_from_fasta(fa, ks)
        dt = time.perf_counter() - t
        size = sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(ex.output_dir) for f in fs)
        print(f"compress={compress} n={n} k={ks}: {dt:.2f} s, {size/1e6:.1f} MB written, {n/dt/1e6:.2f} Mbp/s", flush=True)
