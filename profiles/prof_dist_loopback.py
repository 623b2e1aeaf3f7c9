"""Times (and, under ncu, profiles) the one-launch distributed banded apply on ONE GPU in loopback mode
(JETS_B200_DIST_LOOPBACK=1: the rank is its own neighbour, the operator block-circulant), next to the same rows
without neighbours.  Usage: python profiles/prof_dist_loopback.py [nblk] [block_len] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import jets_b200 as B

nblk = int(sys.argv[1]) if len(sys.argv) > 1 else 32
blk = int(sys.argv[2]) if len(sys.argv) > 2 else 15_625_000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
T = np.float32
B.init(0)
sp = B.JetSpace(T, blk)
own = B.JetBSpace([sp] * nblk)
W = B.rand(own, seed=1)
Sup, Slo, Z = B.JopStencil(T, blk, "fdiff"), B.JopStencil(T, blk, "lap"), B.JopZeroBlock(sp, sp)


def local_rows(wrap):
    rows = []
    for r in range(nblk):
        row = []
        for j in range(nblk + 2):
            c = j - 1
            if c == r:
                row.append(B.JopDiagonal(B.getblock(W, r + 1)))
            elif c == r + 1 and (c < nblk or wrap):
                row.append(Sup)
            elif c == r - 1 and (c >= 0 or wrap):
                row.append(Slo)
            else:
                row.append(Z)
        rows.append(row)
    return B.blockop(rows)


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    B.sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s = torch.cuda.current_stream()
    import ctypes as C
    B.check(B.lib.jets_stream_set(C.c_void_p(s.cuda_stream)))
    a.record(s)
    for _ in range(reps):
        fn()
    b.record(s)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


x, d, m = B.rand(own, seed=2), B.zeros(own), B.zeros(own)
gb = 3 * nblk * blk * 4 / 1e9
for mode in ("plain", "loopback", "loopback+registered"):
    os.environ["JETS_B200_DIST_LOOPBACK"] = "0" if mode == "plain" else "1"
    op = B.dist.DistOp(B, local_rows(mode != "plain"), halo=1)
    if mode.endswith("registered"):
        op.register(x)
    tf = timed(lambda: op.forward(d, x))
    tt = timed(lambda: op.adjoint(m, d))
    print(f"{mode:19s} nblk={nblk} blk={blk}: forward {tf:.4f} ms ({gb / tf * 1e3:.0f} GB/s)  adjoint {tt:.4f} ms ({gb / tt * 1e3:.0f} GB/s)"
          f"  neighbours={op.info(2)} timeouts={op.gate_timeouts}", flush=True)
    op.close()
