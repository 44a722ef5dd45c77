"""Where the time of one rank of the C4 job goes (run on ONE GPU): the byte-range call of rank r of an N-rank job
on the 3.1 Gbp genome, device time per kernel family and wall time per call, for N = 1, 2, 4, 8.

    python tools/c4_range_cost.py [scale]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import bench_extras as bx
from kmerml_b200 import _lib, engine
from kmerml_b200 import dist as kdist


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    lens = bx.human_lengths(scale)
    fasta = bx.gpu_fasta(torch, dev, lens, 4, "synthetic human-sized")
    ctx = _lib.context(0)
    out = []
    for world in (1, 2, 4, 8):
        for rank in sorted({0, world // 2, world - 1}):
            begin, end = kdist.chunk_ranges(int(fasta.numel()), world)[rank]
            call = lambda: engine.count_dense_range_device(fasta, begin, end, [12], None, True)
            for _ in range(3):
                call()
            torch.cuda.synchronize()
            ctx.profile_enable(True)
            ctx.profile_read(reset=True)
            for _ in range(5):
                call()
            torch.cuda.synchronize()
            prof = ctx.profile_read(reset=True)
            ctx.profile_enable(False)
            n = 10
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            a.record()
            for _ in range(n):
                call()
            b.record()
            t_issue = time.perf_counter() - t0
            torch.cuda.synchronize()
            rec = {"world": world, "rank": rank, "range_mb": (end - begin) / 1e6, "gpu_ms_per_call": a.elapsed_time(b) / n,
                   "host_issue_ms_per_call": t_issue * 1e3 / n,
                   "kernel_ms_per_call": {k: prof[k] / 5 for k in
                                          ("ms_partition", "ms_bucket", "ms_finalize", "ms_cascade", "ms_other", "ms_count")},
                   "launches_per_call": prof["launches"] / 5}
            print(json.dumps(rec), flush=True)
            out.append(rec)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/c4_range_cost.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
