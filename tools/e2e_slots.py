#!/usr/bin/env python
"""Host-buffer call against the pipeline depth (KMERML_HOST_SLOTS, read at context creation: one process each).
    python tools/e2e_slots.py [n_genomes] [compact|u32]"""
import os, subprocess, sys
n = sys.argv[1] if len(sys.argv) > 1 else "60"
if os.environ.get("KM_CHILD"):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import time
    import numpy as np
    import torch
    import bench
    from kmerml_b200 import _lib, engine
    dev = torch.device("cuda", 0)
    ks = list(range(1, 13))
    hosts, nb = [], 0
    for i in range(int(n)):
        t, b = bench.make_genome_gpu(i, 1.0, dev, torch)
        h = torch.empty(t.numel(), dtype=torch.uint8, pin_memory=True); h.copy_(t); hosts.append(h); nb += b
    _, row_len = engine.row_layout(ks)
    karr = np.asarray(ks, dtype=np.int32)
    rb = int(_lib.load().kmerml_compact_row_bytes(karr.ctypes.data, len(ks)))
    mode = sys.argv[2] if len(sys.argv) > 2 else "compact"
    rows = torch.empty((int(n), rb), dtype=torch.uint8, pin_memory=True) if mode == "compact" else None
    wide = torch.empty((int(n), row_len), dtype=torch.int32, pin_memory=True) if mode != "compact" else None
    ht = torch.zeros((int(n), len(ks)), dtype=torch.int64, pin_memory=True)
    freq = torch.empty((int(n), row_len), dtype=torch.float32, device=dev)
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        if mode == "compact":
            engine.count_dense_host(hosts, ks, device=dev, out_rows=rows, out_freq=freq, out_totals=ht, compact=True)
        else:
            engine.count_dense_host(hosts, ks, device=dev, out_counts=wide, out_freq=freq, out_totals=ht, freq_on_device=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{mode} slots={os.environ.get('KMERML_HOST_SLOTS')}: {dt * 1e3:7.1f} ms  {nb / dt / 1e9:5.1f} Gbp/s", flush=True)
else:
    for slots in ("2", "3", "4", "6", "8"):
        subprocess.run([sys.executable, __file__, n] + sys.argv[2:3], env=dict(os.environ, KM_CHILD="1", KMERML_HOST_SLOTS=slots))
