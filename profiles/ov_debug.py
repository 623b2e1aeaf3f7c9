"""Multi-GPU overlap diagnostics (torchrun, N>=2): times the halo exchange alone, the interior apply
alone and the overlapped forward/adjoint of dist.OverlappedBanded, on the legacy default stream and on
a private non-blocking stream."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jets_b200 as B  # noqa: E402

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
B.init(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B.dist.init_nccl_from_torch(B, dist, torch, rank, world)
NBLK, BLK, T = int(os.environ.get("OV_NBLK", "32")), 15_625_000, np.float32
part = B.dist.RowPartition(NBLK, world, rank, halo=1)
sp = B.JetSpace(T, BLK)
W = B.rand(B.JetBSpace([sp] * part.nloc), seed=1)
Z = B.JopZeroBlock(sp, sp)
Sup, Slo = B.JopStencil(T, BLK, "fdiff"), B.JopStencil(T, BLK, "lap")


def make_block(r, c):
    if r == c:
        return B.JopDiagonal(B.getblock(W, r - part.r0 + 1))
    return Sup if c == r + 1 else Slo


A = B.dist.build_local_operator(B, part, make_block, lambda: Z)
comm = B.dist.LibComm(B, part)
x, d, m = B.zeros(B.domain(A)), B.zeros(B.range_(A)), B.zeros(B.domain(A))
if os.environ.get("OV_REGISTER", "1") == "1":
    comm.register(x)
    comm.register(m)
ov = B.dist.OverlappedBanded(B, part, comm, make_block, lambda: Z, x, d, m, comm.view)


def timeit(fn, stream, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(n):
        fn()
    b.record(stream)
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(t.item(), 4)


def interior():
    Ai, dv, xv = ov.f_int
    B.mul_(dv, Ai, xv)


def exchange_pair():
    comm.halo_exchange_begin(x, 1, part.nloc)
    comm.halo_exchange_end()


ta = torch.empty(256 * 1024 * 1024, dtype=torch.float32, device="cuda").uniform_()
tb = torch.empty_like(ta)


res = {}
for name, stream in (("private", torch.cuda.Stream()),):
    B.check(B.lib.jets_stream_set(C.c_void_p(stream.cuda_stream)))
    with torch.cuda.stream(stream):
        res[name] = {
            "exchange_only_ms": timeit(lambda: comm.halo_exchange(x, 1, part.nloc), stream),
            "interior_only_ms": timeit(interior, stream),
            "monolithic_fwd_ms": timeit(lambda: B.dist.forward(B, part, comm, A, x, d), stream),
            "overlapped_fwd_ms": timeit(ov.forward, stream),
            "overlapped_adj_ms": timeit(ov.adjoint, stream),
            "exchange_begin_end_ms": timeit(exchange_pair, stream),
        }
if rank == 0:
    print(json.dumps({"world": world, **res}))
B.check(B.lib.jets_dist_shutdown())
dist.destroy_process_group()
