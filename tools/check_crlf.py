"""CRLF (Windows) line ends: parity with the oracle and timing against LF."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import oracle
from kmerml_b200 import engine
rng = np.random.default_rng(5)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12_000_000
seq = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)]
rows = [seq[i:i + 80].tobytes() for i in range(0, n, 80)]
for name, nl in (("LF", b"\n"), ("CRLF", b"\r\n")):
    data = b">chr1 x" + nl + nl.join(rows) + nl
    dev = torch.from_numpy(np.frombuffer(data, np.uint8).copy()).cuda()
    for ks in ([12], [8], [6]):
        engine.count_dense_device(dev, [0, dev.numel()], ks, want_freq=False)
        torch.cuda.synchronize(); t = time.perf_counter()
        res = engine.count_dense_device(dev, [0, dev.numel()], ks, want_freq=False)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
        print(f"{name} dense k={ks}: {dt*1e3:.2f} ms", flush=True)
    if n <= 20_000_000:
        got = (res.counts_of(0, 6).cpu().numpy().astype(np.int64) & 0xFFFFFFFF)
        assert np.array_equal(got, oracle.count_dense(data, 6).astype(np.int64)), "k=6 parity"
        print(name, "parity ok")
