"""C3's distance stage on one GPU: full symmetric call vs the row-block call over all rows vs the plane route.
    python tools/time_distance.py [n] [m]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kmerml_b200 import engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
g = torch.Generator(device="cuda").manual_seed(1)
X = (torch.rand((n, m), device="cuda", generator=g) * 150).to(torch.int32)
X[:, ::97] += 300


def timed(f, reps=10):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        f()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def planes_route():
    p, s, mx = engine.count_planes_device(X)
    nd = engine.planes_needed(int(mx.item()))
    return engine.distance_rows_planes_device(p, nd, s, 0, n)


print(f"pairwise_distance_device      {timed(lambda: engine.pairwise_distance_device(X)):.3f} ms")
print(f"pairwise_distance_rows(0, n)  {timed(lambda: engine.pairwise_distance_rows_device(X, 0, n)):.3f} ms")
print(f"count_planes + rows_planes    {timed(planes_route):.3f} ms")
print(f"count_planes alone            {timed(lambda: engine.count_planes_device(X)):.3f} ms")
p, s, mx = engine.count_planes_device(X)
print(f"rows_planes alone (2 planes)  {timed(lambda: engine.distance_rows_planes_device(p, 2, s, 0, n)):.3f} ms")
print(f"rows_planes 1/8 of the rows   {timed(lambda: engine.distance_rows_planes_device(p, 2, s, 0, n // 8)):.3f} ms")
assert torch.equal(engine.pairwise_distance_device(X), planes_route())
