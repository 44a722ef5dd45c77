#!/usr/bin/env python
"""Wall time of the drop-in KmerFeatureExtractor (statistics CSV) after KmerExtractor (development tool)."""
import os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmerml_b200.kmers.generate import KmerExtractor
from kmerml_b200.kmers.statistics import KmerFeatureExtractor

n = int(sys.argv[1]) if len(sys.argv) > 1 else 12_000_000
ks = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8, 9, 10]
with tempfile.TemporaryDirectory() as tmp:
    fa = os.path.join(tmp, "GCA_000001_synthetic.fna")
    rng = np.random.default_rng(3)
    seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)]
    rows = [seq[i:i + 80].tobytes() for i in range(0, n, 80)]
    with open(fa, "wb") as f:
        f.write(b">chr1 synthetic\n" + b"\n".join(rows) + b"\n")
    ex = KmerExtractor(output_dir=os.path.join(tmp, "kmers"), compress=False)
    t = time.perf_counter()
    org = ex.extract_kmers_from_fasta(fa, ks)
    t1 = time.perf_counter() - t
    fx = KmerFeatureExtractor(input_paths=[os.path.join(tmp, "kmers")], output_dir=os.path.join(tmp, "features"))
    t = time.perf_counter()
    out = fx.extract_features()
    t2 = time.perf_counter() - t
    size = sum(os.path.getsize(os.path.join(dp, f)) for dp, _, fs in os.walk(os.path.join(tmp, "features")) for f in fs)
    rows_n = sum(min(4 ** k, n) for k in ks)
    print(f"extract k={ks}: {t1:.2f} s; features CSV: {t2:.2f} s, {size/1e6:.1f} MB, ~{rows_n} rows -> {t2/rows_n*1e6:.2f} us/row "
          f"(reference: 68-75 us/row)", flush=True)
