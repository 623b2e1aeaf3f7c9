"""Why is the adjoint of bench.py's halo-extended local operator slower than the plain block-tridiagonal
one?  Prints the planner's bundle statistics and kernel times for both structures."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["JETS_B200_PLAN_DEBUG"] = "1"
import jets_b200 as B
import ctypes as C
B.init(0)
stream = torch.cuda.current_stream()
B.check(B.lib.jets_stream_set(C.c_void_p(stream.cuda_stream)))
T = np.float32
nb, blk = int(os.environ.get("NB", "256")), 15_625_000
sp = B.JetSpace(T, blk)
W = B.zeros(B.JetBSpace([sp] * nb))
B.check(B.lib.jets_buf_rand(W._h, 5001, 0, 0))
Z = B.JopZeroBlock(sp, sp)
Sup, Slo = B.JopStencil(T, blk, "fdiff"), B.JopStencil(T, blk, "lap")


def timeit(fn, n=6):
    fn(); fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


for name in ("plain", "ext"):
    if name == "plain":
        A = B.blockop([[B.JopDiagonal(B.getblock(W, r + 1)) if r == c else Sup if c == r + 1 else Slo if c == r - 1 else Z
                        for c in range(nb)] for r in range(nb)])
    else:   # bench.py's local operator: nb x (nb+2) over [lo halo | own | hi halo]
        def blkf(r, j):
            c = j - 1
            if c < 0 or c >= nb:
                return Z
            return B.JopDiagonal(B.getblock(W, r + 1)) if r == c else Sup if c == r + 1 else Slo if c == r - 1 else Z
        A = B.blockop([[blkf(r, j) for j in range(nb + 2)] for r in range(nb)])
    At = B.adjoint(A)
    m = B.rand(B.domain(A), seed=7)
    d = B.zeros(B.range_(A))
    m2 = B.zeros(B.domain(A))
    print(name, "forward plan:", file=sys.stderr); B.mul_(d, A, m)
    print(name, "adjoint plan:", file=sys.stderr); B.mul_(m2, At, d)
    f, t = timeit(lambda: B.mul_(d, A, m)), timeit(lambda: B.mul_(m2, At, d))
    both = timeit(lambda: (B.mul_(d, A, m), B.mul_(m2, At, d)))
    print(f"{name}: fwd {f:.3f} ms adj {t:.3f} ms  fwd+adj {both:.3f} ms", flush=True)
    del A, At, m, d, m2
