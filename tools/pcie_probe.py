"""What the link of this box gives the host call: pinned H2D in genome-sized pieces, D2H in wire-row-sized pieces,
and both at once (two streams) -- the ceiling of `e2e` (2.70 GB up + 1.47 GB down per C2 step).
    python tools/pcie_probe.py"""
import json
import torch

dev = torch.device("cuda:0")
n_up, piece_up = 100, 27_000_000
n_dn, piece_dn = 100, 14_700_000
hu = torch.empty(n_up * piece_up, dtype=torch.uint8).pin_memory()
du = torch.empty(n_up * piece_up, dtype=torch.uint8, device=dev)
hd = torch.empty(n_dn * piece_dn, dtype=torch.uint8).pin_memory()
dd = torch.empty(n_dn * piece_dn, dtype=torch.uint8, device=dev)
hu.fill_(65); dd.fill_(1)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def up():
    with torch.cuda.stream(s1):
        for i in range(n_up):
            du[i * piece_up:(i + 1) * piece_up].copy_(hu[i * piece_up:(i + 1) * piece_up], non_blocking=True)


def down():
    with torch.cuda.stream(s2):
        for i in range(n_dn):
            hd[i * piece_dn:(i + 1) * piece_dn].copy_(dd[i * piece_dn:(i + 1) * piece_dn], non_blocking=True)


def timed(fns, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        s1.wait_stream(torch.cuda.current_stream()); s2.wait_stream(torch.cuda.current_stream())
        for f in fns:
            f()
        torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


out = {}
ms = timed([up]); out["h2d_alone"] = {"ms": ms, "GB/s": n_up * piece_up / ms / 1e6}
ms = timed([down]); out["d2h_alone"] = {"ms": ms, "GB/s": n_dn * piece_dn / ms / 1e6}
ms = timed([up, down]); out["both"] = {"ms": ms, "h2d_GB/s": n_up * piece_up / ms / 1e6, "d2h_GB/s": n_dn * piece_dn / ms / 1e6,
                                        "e2e_ceiling_Gbp/s": 2.664 / (ms * 1e-3)}
print(json.dumps(out, indent=1))
