"""Profiling driver (run under ncu on the GPU box): a few launches of the fused engine on
config-5-like (block-tridiagonal, 3 terms/row), config-2 (3-stage chain) and config-1 shapes."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jets_b200 as B

which = sys.argv[1] if len(sys.argv) > 1 else "c5"
T = np.float32
B.init(0)
B.set_fused_engine(os.environ.get("PROF_ENGINE", "auto"))   # auto | tma_nocache | ldg
if which == "c5":
    nb, blk = 32, 3_906_252
    sp = B.JetSpace(T, blk)
    W = B.rand(B.JetBSpace([sp] * nb), seed=1)
    Z = B.JopZeroBlock(sp, sp)
    Su, Sl = B.JopStencil(T, blk, "fdiff"), B.JopStencil(T, blk, "lap")
    A = B.blockop([[B.JopDiagonal(B.getblock(W, r + 1)) if r == c else Su if c == r + 1 else Sl if c == r - 1 else Z
                    for c in range(nb)] for r in range(nb)])
elif which == "c2":
    n = 100_000_000
    sp = B.JetSpace(T, n)
    A = B.jacobian(B.JopDiagonal(B.rand(sp, seed=1)) @ B.JopStencil(T, n, "fdiff") @ B.JopPointwise(T, n, "square"),
                   B.rand(sp, seed=2))
elif which == "c4":
    nb, n, T = 8, 1 << 20, np.float64
    sp = B.JetSpace(T, n)
    W = B.rand(B.JetBSpace([sp] * nb), seed=4001)
    Bd = B.blockop([[B.JopDiagonal(B.getblock(W, i + 1)) if i == j else B.JopZeroBlock(sp, sp) for j in range(nb)] for i in range(nb)])
    Sd = B.blockop([[B.JopStencil(T, n, "lap") if i == j else B.JopZeroBlock(sp, sp) for j in range(nb)] for i in range(nb)])
    A = Bd - 0.5 * Sd
else:
    n = 1_000_000
    sp = B.JetSpace(np.float64, n)
    W = B.rand(B.JetBSpace([sp] * 16), seed=1)
    A = B.blockop([[B.JopDiagonal(B.getblock(W, 1 + r + 4 * c)) for c in range(4)] for r in range(4)])
m = B.rand(B.domain(A), seed=3)
d = B.zeros(B.range_(A))
m2 = B.zeros(B.domain(A))
for _ in range(3):
    B.mul_(d, A, m)
    B.mul_(m2, A.T, d)
B.sync()
print(which, B.plan_info(A))
