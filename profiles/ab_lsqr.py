"""Config 4 (LSQR, 200 iterations, A = B - 0.5 S on 8 x 2^20 Float64) per-iteration time of the three solver variants;
env switches (JETS_B200_NO_FUSED_NORM, JETS_B200_NO_PRE_STATE ...) are read at jets_init: run once per setting."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, ctypes as C
import jets_b200 as B
B.init(0)
s4 = torch.cuda.Stream()
B.check(B.lib.jets_stream_set(C.c_void_p(s4.cuda_stream)))
nb, n4, iters = 8, 1 << 20, 200
T8 = np.float64
sp = B.JetSpace(T8, n4)
W4 = B.rand(B.JetBSpace([sp] * nb), seed=4001)
Bd = B.blockop([[B.JopDiagonal(B.getblock(W4, i + 1)) if i == j else B.JopZeroBlock(sp, sp) for j in range(nb)] for i in range(nb)])
Sd = B.blockop([[B.JopStencil(T8, n4, "lap") if i == j else B.JopZeroBlock(sp, sp) for j in range(nb)] for i in range(nb)])
A4 = Bd - 0.5 * Sd
rhs4 = B.rand(B.range_(A4), seed=4002)
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("JETS_B200_")) or "default"
only = os.environ.get("AB_LSQR_ONLY", "unfused,fused").split(",")
iters = int(os.environ.get("AB_LSQR_ITERS", iters))
for name, cls in (("unfused", B.solvers.LsqrGraph), ("fused", B.solvers.LsqrGraphFused)):
    if name not in only:
        continue
    G = cls(A4, rhs4)
    G.run(5)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s4)
        G.run(iters)
        e1.record(s4)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / iters)
    x, (a, b) = G.result()
    print(f"{tag}: {name:8s} {best:7.2f} us/iteration  ({16 * nb * n4 * 8 / best / 1e3:.0f} GB/s of the 16 N w the fused form needs)  |x| = {float(B.norm(x)):.12g}", flush=True)
