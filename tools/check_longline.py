"""Unwrapped (single-line) FASTA: parity with the oracle and timing against the 80-column layout."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import oracle
from kmerml_b200 import engine

rng = np.random.default_rng(5)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12_000_000
seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)]
def wrap(width):
    if width == 0:
        return b">chr1 unwrapped\n" + seq.tobytes() + b"\n>chr2\n" + seq[: n // 3].tobytes() + b"\n"
    rows = [seq[i:i + width].tobytes() for i in range(0, n, width)]
    return b">chr1 wrapped\n" + b"\n".join(rows) + b"\n>chr2\n" + b"\n".join(rows[: len(rows) // 3]) + b"\n"
for width in (80, 0):
    data = wrap(width)
    dev = torch.from_numpy(np.frombuffer(data, np.uint8).copy()).cuda()
    for ks in ([12], [8], [6]):
        engine.count_dense_device(dev, [0, dev.numel()], ks, want_freq=False)
        torch.cuda.synchronize(); t = time.perf_counter()
        res = engine.count_dense_device(dev, [0, dev.numel()], ks, want_freq=False)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        print(f"width={width} dense k={ks}: {dt*1e3:.2f} ms", flush=True)
    if width == 0 and n <= 20_000_000:
        want = oracle.count_dense(data, 8)
        got = (res.counts_of(0, 6).cpu().numpy().astype(np.int64) & 0xFFFFFFFF)
        assert np.array_equal(got, oracle.count_dense(data, 6)), "k=6 parity"
    t = time.perf_counter()
    st = engine.genome_stats_device(dev); torch.cuda.synchronize()
    print(f"width={width} genome_stats: {(time.perf_counter()-t)*1e3:.2f} ms {st}", flush=True)
    t = time.perf_counter()
    out = engine.count_sparse_device(dev, 20); torch.cuda.synchronize()
    print(f"width={width} sparse k=20: {(time.perf_counter()-t)*1e3:.2f} ms unique={out[0].numel()}", flush=True)
