#!/usr/bin/env python
"""Low-complexity genomes (tandem repeats overflow the partition path's slots): parity and timing."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import oracle
from kmerml_b200 import engine
rng = np.random.default_rng(9)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 26_000_000
def genome(frac_repeat, unit_len):
    seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)].copy()
    unit = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, unit_len)]
    nrep = int(n * frac_repeat)
    block = 200_000
    for a in range(0, nrep, block):                       # repeat blocks spread over the genome
        pos = int(a / max(frac_repeat, 1e-9)) if frac_repeat else 0
        pos = min(pos, n - block)
        seq[pos:pos + block] = np.resize(unit, block)
    rows = [seq[i:i + 80].tobytes() for i in range(0, n, 80)]
    return b">chr1\n" + b"\n".join(rows) + b"\n"
for name, frac, unit in (("random", 0.0, 1), ("10% 171-bp satellite", 0.10, 171), ("10% dinucleotide (AC)n", 0.10, 2),
                         ("100% 171-bp satellite", 1.0, 171)):
    data = genome(frac, unit)
    dev = torch.from_numpy(np.frombuffer(data, np.uint8).copy()).cuda()
    line = []
    for ks in (list(range(1, 13)), [8], [6]):
        engine.count_dense_device(dev, [0, dev.numel()], ks)
        torch.cuda.synchronize(); t = time.perf_counter()
        res = engine.count_dense_device(dev, [0, dev.numel()], ks)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        ok = ""
        if n <= 30_000_000:
            kk = max(ks)
            got = res.counts_of(0, kk).cpu().numpy().astype(np.int64) & 0xFFFFFFFF
            ok = "ok" if np.array_equal(got, oracle.count_dense(data, kk).astype(np.int64)) else "MISMATCH"
        line.append(f"k={'1..12' if len(ks) > 1 else ks[0]} {dt*1e3:.2f} ms {ok}")
    print(f"{name}: " + "   ".join(line), flush=True)
