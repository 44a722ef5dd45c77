"""Multi-rank host logic on CPU: world_size-2 gloo runs of kmerml_b200.dist with the CUDA entry
points replaced by the CPU thread emulator (tests/emu), checked against the oracle.  No GPU."""
import ctypes
import os
import socket
import subprocess

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from kmerml_b200 import dist as kdist
from kmerml_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))


def _emu():
    src = os.path.join(HERE, "emu", "emu_dense.cpp")
    so = os.path.join(HERE, "emu", "libemu_dense.so")
    hdr = os.path.join(HERE, "..", "kmerml_b200", "csrc", "fasta_walk.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-o", so, src])
    L = ctypes.CDLL(so)
    L.emu_count_dense_range.restype = ctypes.c_int64
    L.emu_count_dense_range.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p]
    return L


def emu_count_range(fasta, begin, end, ks, min_record_len, canonical):
    """Stand-in for engine.count_dense_range_device with the same contract (CPU emulator)."""
    a = fasta.numpy()
    kmax = max(ks)
    levels = np.zeros(sum(4 ** j for j in range(1, kmax + 1)), np.uint64)
    if begin < end:
        _emu().emu_count_dense_range(a.ctypes.data, a.size, kmax, min_record_len or kmax, 512, 1, begin, end,
                                     levels.ctypes.data)
    offs = np.concatenate(([0], np.cumsum([4 ** j for j in range(1, kmax + 1)])))
    row = np.concatenate([levels[offs[k - 1]:offs[k]] for k in ks])
    totals = np.array([levels[offs[k - 1]:offs[k]].sum() for k in ks], dtype=np.int64)
    return torch.from_numpy(row.astype(np.uint32).view(np.int32).copy()), torch.from_numpy(totals)


def emu_count_batch(fastas, ks, min_record_len, canonical):
    rows, tots = [], []
    for f in fastas:
        c, t = emu_count_range(f, 0, int(f.numel()), ks, min_record_len, canonical)
        rows.append(c)
        tots.append(t)
    _, row_len = __import__("kmerml_b200.engine", fromlist=["row_layout"]).row_layout(ks)
    if not rows:
        return torch.zeros((0, row_len), dtype=torch.int32), torch.zeros((0, len(ks)), dtype=torch.int64)
    return torch.stack(rows), torch.stack(tots)


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ks = [3, 8]
        # (1) one genome cut into byte ranges + all-reduce
        genome = synth.fasta_bytes([30_000, 7, 21_000], seed=5)
        genome[40_000:40_300] = ord("N")
        data = torch.from_numpy(genome.copy())
        counts, freq, totals = kdist.count_genome_chunked(data, ks, count_range=emu_count_range)
        ok = True
        off = 0
        for ki, k in enumerate(ks):
            ref = oracle.count_dense(genome.tobytes(), k, max(ks))
            got = counts[off:off + 4 ** k].numpy().view(np.uint32).astype(np.uint64)
            ok &= bool(np.array_equal(ref, got)) and int(totals[ki]) == int(ref.sum())
            f = freq[off:off + 4 ** k].numpy().astype(np.float64)
            ok &= bool(np.all(np.abs(f - oracle.frequencies(ref)) <= 1e-6 * np.maximum(oracle.frequencies(ref), 1e-300)))
            off += 4 ** k
        # (2) genomes sharded across ranks + gather
        gs = [synth.fasta_bytes([3000 + 700 * i, 1500], seed=20 + i) for i in range(5)]
        idx, c2, t2 = kdist.count_genomes_sharded([torch.from_numpy(g.copy()) for g in gs], ks, count_batch=emu_count_batch)
        ok &= idx == list(range(5))
        for i, g in enumerate(gs):
            off = 0
            for ki, k in enumerate(ks):
                ref = oracle.count_dense(g.tobytes(), k, max(ks))
                ok &= bool(np.array_equal(ref, c2[i, off:off + 4 ** k].numpy().view(np.uint32).astype(np.uint64)))
                ok &= int(t2[i, ki]) == int(ref.sum())
                off += 4 ** k
        # (3) one genome, sparse k: byte ranges -> all-to-all by key range -> merge
        from helpers import reduce_windows, unwrapped_fasta_with_windows
        rng = np.random.default_rng(77)
        seqs = [np.frombuffer(b"ACGTN", np.uint8)[rng.choice(5, n, p=[.24, .25, .25, .24, .02])].tobytes()
                for n in (300_000, 10, 150_000)]
        for k, canonical in ((17, False), (9, True)):
            fa, wk, we = unwrapped_fasta_with_windows(seqs, k, canonical)

            def np_count_range(fasta, begin, end, kk, ml, c):
                u, cnt, fst = reduce_windows(wk, we, begin, end)
                return (torch.from_numpy(u.view(np.int64).copy()), torch.from_numpy(cnt.astype(np.int32)),
                        torch.from_numpy(fst.astype(np.int32)), int(cnt.sum()))

            def np_merge(keys, counts, first, kk):
                kn, cn, fn = keys.numpy().view(np.uint64), counts.numpy().astype(np.int64), first.numpy().astype(np.int64)
                order = np.lexsort((fn, kn))
                kn, cn, fn = kn[order], cn[order], fn[order]
                u, idx = np.unique(kn, return_index=True)
                return (torch.from_numpy(u.view(np.int64).copy()), torch.from_numpy(np.add.reduceat(cn, idx).astype(np.int32)),
                        torch.from_numpy(fn[idx].astype(np.int32)))

            mk, mc, mf, windows = kdist.count_sparse_sharded(torch.from_numpy(np.frombuffer(fa, np.uint8).copy()), k,
                                                             canonical=canonical, count_range=np_count_range, merge=np_merge)
            # ... and the RAW route: windows grouped by owner, exchanged, sorted + reduced once
            def np_emit_range(fasta, begin, end, kk, owner_bits, ml, c):
                sel = (we >= begin) & (we < end)
                kk_, ee_ = wk[sel], we[sel]
                owner = (kk_ >> np.uint64(2 * kk - owner_bits)).astype(np.int64) if owner_bits else np.zeros(kk_.size, np.int64)
                order = np.argsort(owner, kind="stable")
                return (torch.from_numpy(kk_[order].view(np.int64).copy()), torch.from_numpy(ee_[order].astype(np.int32)),
                        np.bincount(owner, minlength=1 << owner_bits).tolist())

            def np_reduce_windows(keys, ends, sort_bits):
                kn, en = keys.numpy().view(np.uint64), ends.numpy().astype(np.int64)
                order = np.lexsort((en, kn))
                kn, en = kn[order], en[order]
                uu, idx, cc = np.unique(kn, return_index=True, return_counts=True)
                return (torch.from_numpy(uu.view(np.int64).copy()), torch.from_numpy(cc.astype(np.int32)),
                        torch.from_numpy(en[idx].astype(np.int32)))

            rk, rc_, rf, rwin = kdist.count_sparse_sharded(torch.from_numpy(np.frombuffer(fa, np.uint8).copy()), k,
                                                          canonical=canonical, emit_range=np_emit_range,
                                                          reduce_windows=np_reduce_windows)
            ok &= bool(torch.equal(rk, mk) and torch.equal(rc_, mc) and torch.equal(rf, mf)) and rwin == windows
            u, cnt, fst = reduce_windows(wk, we)
            ok &= windows == int(cnt.sum())
            # this rank holds exactly the k-mers of its key range; the ranges tile the key space in rank order
            top = (u >> np.uint64(2 * k - 16)) & np.uint64(0xFFFF)
            mine = ((top * np.uint64(world)) >> np.uint64(16)) == rank
            ok &= bool(np.array_equal(mk.numpy().view(np.uint64), u[mine]))
            ok &= bool(np.array_equal(mc.numpy().astype(np.int64), cnt[mine]))
            ok &= bool(np.array_equal(mf.numpy().astype(np.int64), fst[mine]))
        # (4) one genome, dense: byte ranges -> reduce-scatter to owner slices (canonical fold before the
        # reduction: it is linear), the slices concatenated in rank order give the single-GPU row
        def emu_canonical_range(fasta, begin, end, kk, ml, canonical):
            c, t = emu_count_range(fasta, begin, end, kk, ml, False)
            if canonical:
                fwd = c.numpy().view(np.uint32).astype(np.uint64)
                c = torch.from_numpy(oracle.canonical_from_forward(fwd, kk[0]).astype(np.uint32).view(np.int32).copy())
            return c, t
        for canonical in (False, True):
            part, off, windows = kdist.count_genome_chunked_scatter(data, 6, canonical=canonical, count_range=emu_canonical_range)
            ref = oracle.count_dense(genome.tobytes(), 6, 6)
            if canonical:
                ref = oracle.canonical_from_forward(ref, 6)
            n = 4 ** 6 // world
            ok &= off == rank * n and windows == int(ref.sum())
            ok &= bool(np.array_equal(part.numpy().view(np.uint32).astype(np.uint64), ref[off:off + n]))
        # (5) distance matrix in row blocks + all_gather
        X = torch.from_numpy(np.random.default_rng(5).integers(0, 50, (7, 64)).astype(np.int32))

        def np_rows(counts, r0, r1, metric):
            D = oracle.pairwise_distance(counts.numpy().astype(np.float64), metric)
            return torch.from_numpy(D[r0:r1].astype(np.float32))
        for metric in ("cosine", "euclidean"):
            D = kdist.distance_matrix_sharded(X, metric, rows_fn=np_rows)
            ok &= bool(np.array_equal(D.numpy(), oracle.pairwise_distance(X.numpy().astype(np.float64), metric).astype(np.float32)))
        # (6) the same from sharded rows, gathered as byte planes (uneven shards: 4 + 3 genomes, one count >= 256
        # so that two planes travel); numpy stand-ins for the two kernels
        shards = [[0, 2, 5, 6], [1, 3, 4]]
        X2 = X.clone()
        X2[3, 7] = 70000

        def np_planes(counts, n_rows):
            c = counts.numpy().view(np.uint32)
            planes = np.zeros((4, n_rows, c.shape[1]), np.uint8)
            for d in range(4):
                planes[d, :c.shape[0]] = (c >> (8 * d)) & 0xFF
            sumsq = np.zeros(n_rows)
            sumsq[:c.shape[0]] = (c.astype(np.float64) ** 2).sum(axis=1)
            return torch.from_numpy(planes), torch.from_numpy(sumsq), torch.tensor([int(c.max()) if c.size else 0], dtype=torch.int32)

        def np_rows_planes(planes, nd, sumsq, r0, r1, metric):
            p = planes.numpy().astype(np.float64)
            full = sum(p[d] * 256.0 ** d for d in range(nd))
            with np.errstate(invalid="ignore", divide="ignore"):
                D = oracle.pairwise_distance(full, metric)
            return torch.from_numpy(D[r0:r1].astype(np.float32))
        for metric in ("cosine", "euclidean"):
            info = {}
            D = kdist.distance_matrix_from_shards(X2[shards[rank]], shards, metric, planes_fn=np_planes, rows_fn=np_rows_planes,
                                                  info=info)
            want = oracle.pairwise_distance(X2.numpy().astype(np.float64), metric).astype(np.float32)
            ok &= info["planes"] == 3 and bool(np.array_equal(D.numpy(), want))
        results[rank] = ok
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_rank_gloo_chunked_and_sharded():
    _emu()
    oracle.build()
    world = 2
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
        assert dict(results) == {0: True, 1: True}


def test_shard_genomes_lpt():
    sizes = [40, 12, 33, 25, 18, 30, 12]
    shards = kdist.shard_genomes(sizes, 3)
    assert sorted(i for s in shards for i in s) == list(range(7))
    loads = [sum(sizes[i] for i in s) for s in shards]
    assert max(loads) - min(loads) <= max(sizes)
    assert shards == kdist.shard_genomes(sizes, 3)
    assert kdist.shard_genomes([5, 5], 4) == [[0], [1], [], []]


def test_chunk_ranges_tile_aligned():
    for n in (0, 1, 16384, 16385, 1_000_000, 3_100_000_123):
        for w in (1, 2, 4, 8):
            r = kdist.chunk_ranges(n, w)
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert all(b % kdist.TILE_BYTES == 0 for b, _ in r if b < n)


def test_single_process_additivity_over_ranges():
    """Counts of byte ranges that tile the file add up to the whole (the multi-GPU invariant)."""
    genome = synth.fasta_bytes([50_000, 12_345], seed=9)
    data = torch.from_numpy(genome.copy())
    for world in (2, 3, 5):
        total = None
        for b, e in kdist.chunk_ranges(genome.size, world):
            c, _ = emu_count_range(data, b, e, [2, 9], 9, False)
            total = c.numpy().view(np.uint32).astype(np.uint64) if total is None else total + c.numpy().view(np.uint32)
        ref = np.concatenate([oracle.count_dense(genome.tobytes(), k, 9) for k in (2, 9)])
        assert np.array_equal(total, ref), world


def test_key_owner_splits_tile_the_key_space():
    """Owners are ascending along sorted keys, every key has exactly one owner, also for k = 32 (top bit set)."""
    rng = np.random.default_rng(3)
    for k in (5, 8, 21, 31, 32):
        bits = 2 * k
        keys = np.unique(rng.integers(0, 2 ** 63, 5000, dtype=np.uint64) >> np.uint64(64 - bits)) if bits < 64 else \
            np.unique(rng.integers(0, 2 ** 63, 5000, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, 5000, dtype=np.uint64))
        keys.sort()
        t = torch.from_numpy(keys.view(np.int64).copy())
        for world in (1, 2, 3, 8):
            sp = kdist.key_owner_splits(t, k, world)
            assert len(sp) == world and sum(sp) == keys.size
            top = (keys >> np.uint64(bits - 16)) & np.uint64(0xFFFF) if bits >= 16 else (keys << np.uint64(16 - bits)) & np.uint64(0xFFFF)
            owner = (top * np.uint64(world)) >> np.uint64(16)
            assert np.all(np.diff(owner.astype(np.int64)) >= 0)
            assert sp == np.bincount(owner.astype(np.int64), minlength=world).tolist()
    assert kdist.key_owner_splits(torch.zeros(0, dtype=torch.int64), 21, 4) == [0, 0, 0, 0]
