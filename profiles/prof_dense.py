"""Profiling driver (run under ncu on the GPU box): dense JopBlock GEMV, both orientations, on a
16x16 grid of 2048x2048 Float32 blocks (4 GiB of matrices >> L2; config 3 at 1/16 size)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jets_b200 as B

nb, k = 16, 2048
T = np.float32
nrhs = int(sys.argv[1]) if len(sys.argv) > 1 else 1
mats = B.zeros(B.JetBSpace([B.JetSpace(T, k, k)] * (nb * nb)))
B.check(B.lib.jets_buf_rand(mats._h, 3001, 0, 0))
A = B.blockop([[B.JopDense(B.getblock(mats, 1 + r + nb * c), nrhs=nrhs) for c in range(nb)] for r in range(nb)])
m = B.rand(B.domain(A), seed=3)
d = B.zeros(B.range_(A))
m2 = B.zeros(B.domain(A))
import time
At = A.T
for _ in range(3):
    B.mul_(d, A, m)
    B.mul_(m2, At, d)
B.sync()
for nm, fn in (("fwd", lambda: B.mul_(d, A, m)), ("adj", lambda: B.mul_(m2, At, d))):
    B.sync()
    t0 = time.perf_counter()
    for _ in range(10):
        fn()
    B.sync()
    ms = (time.perf_counter() - t0) * 100
    print(f"dense nrhs={nrhs} {nm}: {ms:.3f} ms  {nb * nb * k * k * 4 / ms / 1e6:.0f} GB/s of matrix bytes")
print("dense", nrhs, B.plan_info(A))
