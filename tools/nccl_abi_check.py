"""kmerml_allreduce_counts against a plain sum: one process per GPU (torchrun), a ncclComm_t made with NCCL's own
API through ctypes (the unique id travels over a gloo group), the C4 exchange through the C-ABI, the expected row
from a gloo all-reduce of the same data on the CPU.  Also runs as a single process (a communicator of one rank).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/nccl_abi_check.py
"""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch
import torch.distributed as dist

from kmerml_b200 import _lib


class NcclUniqueId(ctypes.Structure):
    _fields_ = [("internal", ctypes.c_byte * 128)]


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.zeros(1, device=dev)                                   # CUDA context before NCCL
    if world > 1:
        dist.init_process_group("gloo")
    nccl = ctypes.CDLL("libnccl.so.2")                            # torch has loaded its bundled copy already
    uid = NcclUniqueId()
    if rank == 0:
        assert nccl.ncclGetUniqueId(ctypes.byref(uid)) == 0
    if world > 1:
        box = [bytes(uid.internal)]
        dist.broadcast_object_list(box, src=0)
        ctypes.memmove(ctypes.byref(uid), box[0], 128)
    comm = ctypes.c_void_p()
    nccl.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, NcclUniqueId, ctypes.c_int]
    assert nccl.ncclCommInitRank(ctypes.byref(comm), world, uid, rank) == 0
    L, ctx = _lib.load(), _lib.context(local)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    ok = True
    for dtype, np_t, torch_t in ((0, np.uint32, torch.int32), (1, np.uint64, torch.int64)):
        n = 4 ** 8 * world
        rng = np.random.default_rng(100 + rank)
        mine = rng.integers(0, 2 ** 31 if dtype == 0 else 2 ** 40, size=n, dtype=np.int64)
        want = torch.from_numpy(mine.copy())
        if world > 1:
            dist.all_reduce(want)                                 # gloo, int64: the plain sum
        want = want.numpy().astype(np.uint64)
        if dtype == 0:
            want = want & 0xFFFFFFFF                              # uint32 sums wrap
        for scatter in (0, 1):
            d = torch.from_numpy(mine.astype(np_t).view(np.int32 if dtype == 0 else np.int64)).to(dev)
            _lib.check(L.kmerml_allreduce_counts(ctx.handle, comm, d.data_ptr(), n, dtype, scatter, stream))
            torch.cuda.synchronize()
            got = d.cpu().numpy().view(np_t).astype(np.uint64)
            if scatter:
                per = n // world
                ok &= bool(np.array_equal(got[rank * per:(rank + 1) * per], want[rank * per:(rank + 1) * per]))
            else:
                ok &= bool(np.array_equal(got, want))
    # argument errors come back as codes, not crashes
    ok &= L.kmerml_allreduce_counts(ctx.handle, None, None, 0, 0, 0, stream) < 0
    if world > 1:
        ok &= L.kmerml_allreduce_counts(ctx.handle, comm, torch.zeros(world + 1, dtype=torch.int32, device=dev).data_ptr(),
                                        world + 1, 0, 1, stream) < 0
    nccl.ncclCommDestroy.argtypes = [ctypes.c_void_p]
    nccl.ncclCommDestroy(comm)
    if world > 1:
        flags = [None] * world
        dist.all_gather_object(flags, ok)
        ok = all(flags)
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps({"kmerml_allreduce_counts": "ok" if ok else "MISMATCH", "ranks": world}))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
