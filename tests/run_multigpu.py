#!/usr/bin/env python
"""Multi-GPU check, launched as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/run_multigpu.py
One large synthetic genome is resident on every GPU; rank r counts byte range r, one NCCL
all-reduce sums the dense rows (BASELINE config 4 in miniature), and the result must be
bit-identical to the single-GPU count.  Also exercises genome sharding + all_gather."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmerml_b200 import dist as kdist          # noqa: E402
from kmerml_b200 import engine, synth          # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    mbp = float(os.environ.get("KM_MBP", "120"))
    rng = np.random.default_rng(4)
    lens = synth.split_lengths(int(mbp * 1e6), 24, rng)
    genome = synth.fasta_bytes(lens, seed=4)
    fasta = torch.from_numpy(genome).to(dev)
    ks = [12]
    torch.cuda.synchronize()
    for canonical in (False, True):
        t0 = time.perf_counter()
        counts, freq, totals = kdist.count_genome_chunked(fasta, ks, canonical=canonical)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        whole = engine.count_dense_device(fasta, [0, int(fasta.numel())], ks, canonical=canonical)
        torch.cuda.synchronize()
        same = torch.equal(counts, whole.counts[0]) and torch.equal(totals, whole.totals[0])
        fsame = torch.allclose(freq, whole.freq[0], rtol=1e-6, atol=0)
        flag = torch.tensor([1 if (same and fsame) else 0], device=dev)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"chunked genome {mbp:.0f} Mbp k=12 canonical={canonical} world={world}: "
                  f"bit-exact vs single GPU = {bool(flag.item())}  ({dt * 1e3:.1f} ms incl. all-reduce)")
        assert flag.item() == 1
    # genome sharding + gather
    gs = [torch.from_numpy(synth.config3_genome(i, scale=0.05)).to(dev) for i in range(11)]
    idx, c, t = kdist.count_genomes_sharded(gs, [8])
    buf = torch.cat(gs)
    offs = np.concatenate(([0], np.cumsum([int(g.numel()) for g in gs]))).tolist()
    ref = engine.count_dense_device(buf, offs, [8], want_freq=False)
    ok = idx == list(range(11)) and torch.equal(c, ref.counts) and torch.equal(t, ref.totals)
    if rank == 0:
        print(f"sharded 11 genomes over {world} ranks + all_gather: identical to single GPU = {ok}")
    assert ok
    # one genome with N runs, sparse k = 21 canonical (BASELINE config 5 in miniature): byte ranges -> all-to-all
    # by key range -> merge; the ranks' shards, concatenated in rank order, must equal the single-GPU result
    g5 = genome.copy()
    for a in range(3_000_000, g5.size - 100_000, 9_000_000):
        seg = g5[a:a + 40_000]
        seg[(seg != 10) & (seg != 62)] = ord("N")
    fasta5 = torch.from_numpy(g5).to(dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mk, mc, mf, windows = kdist.count_sparse_sharded(fasta5, 21, canonical=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    wk, wc, wf, ww = engine.count_sparse_device(fasta5, 21, canonical=True)
    if world > 1:
        n_mine = torch.tensor([mk.numel()], dtype=torch.int64, device=dev)
        sizes = [torch.empty_like(n_mine) for _ in range(world)]
        dist.all_gather(sizes, n_mine)
        lo = int(sum(int(x.item()) for x in sizes[:rank]))
        total = int(sum(int(x.item()) for x in sizes))
    else:
        lo, total = 0, mk.numel()
    ok5 = (total == wk.numel() and windows == ww and torch.equal(mk, wk[lo:lo + mk.numel()])
           and torch.equal(mc, wc[lo:lo + mk.numel()]) and torch.equal(mf, wf[lo:lo + mk.numel()]))
    flag = torch.tensor([1 if ok5 else 0], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"sparse k=21 canonical, {mbp:.0f} Mbp with N runs, world={world}: {total} distinct k-mers, shards identical to "
              f"single GPU = {bool(flag.item())}  ({dt * 1e3:.1f} ms incl. all-to-all)")
    assert flag.item() == 1
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
