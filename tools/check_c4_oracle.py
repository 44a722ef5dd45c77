#!/usr/bin/env python
"""BASELINE config 4 at FULL size against the oracle: the bench's 3.1 Gbp synthetic genome (24 records, seed 4),
k = 12, counted on the GPU (forward and canonical) and by the oracle C port on all host cores -- one record per
task (counts of records add: the reference counts per record into one dict, generate.py:36-58) -- bit for bit.
The canonical row is checked through the fold of the oracle's forward row.  ~1-2 minutes of CPU.

    python tools/check_c4_oracle.py [scale]        # writes profiles/r02_c4_oracle_check.json
"""
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_extras                                 # noqa: E402
import oracle                                       # noqa: E402  (checker only)
from kmerml_b200 import engine                      # noqa: E402


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    k = 12
    dev = torch.device("cuda", 0)
    lens = bench_extras.human_lengths(scale)
    fasta = bench_extras.gpu_fasta(torch, dev, lens, 4, "synthetic human-sized")
    n = int(fasta.numel())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fwd = engine.count_dense_device(fasta, [0, n], [k], want_freq=False)
    can = engine.count_dense_device(fasta, [0, n], [k], canonical=True, want_freq=False)
    torch.cuda.synchronize()
    t_gpu = time.perf_counter() - t0
    g_fwd = fwd.counts[0].cpu().numpy().view(np.uint32).astype(np.uint64)
    g_can = can.counts[0].cpu().numpy().view(np.uint32).astype(np.uint64)
    host = fasta.cpu().numpy()
    del fasta
    # records: every header line starts a new file for the oracle
    starts = np.flatnonzero(host == ord(">"))
    bounds = list(starts) + [n]
    oracle.build()
    cores = max(1, min(os.cpu_count() or 1, 32))

    def one(i):
        return oracle.count_dense(host[bounds[i]:bounds[i + 1]].tobytes(), k, k)

    t0 = time.perf_counter()
    total = np.zeros(4 ** k, np.uint64)
    with ThreadPoolExecutor(cores) as ex:
        for c in ex.map(one, range(len(starts))):
            total += c
    t_cpu = time.perf_counter() - t0
    o_can = oracle.canonical_from_forward(total, k)
    res = {
        "workload": f"C4: one {sum(lens) / 1e9:.2f} Gbp synthetic genome, {len(lens)} records, k=12, bench_extras.gpu_fasta(seed 4)",
        "windows": int(total.sum()), "gpu_windows": int(fwd.totals[0, 0]),
        "forward_bit_exact_vs_oracle": bool(np.array_equal(g_fwd, total)),
        "canonical_bit_exact_vs_fold_of_oracle_forward": bool(np.array_equal(g_can, o_can)),
        "oracle_seconds": round(t_cpu, 1), "oracle_threads": cores, "oracle_Gbp_per_s": sum(lens) / t_cpu / 1e9,
        "gpu_seconds_two_counts_incl_allocation": round(t_gpu, 3),
    }
    print(json.dumps(res))
    if scale == 1.0:
        with open(os.path.join(ROOT, "gpurun_out", "r02_c4_oracle_check.json"), "w") as f:
            json.dump(res, f, indent=1)
    assert res["forward_bit_exact_vs_oracle"] and res["canonical_bit_exact_vs_fold_of_oracle_forward"]


if __name__ == "__main__":
    main()
