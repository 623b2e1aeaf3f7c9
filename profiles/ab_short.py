"""A/B of the short-launch configs (C1 4x4 diagonal f64 1e6; C4's operator 8x8 block-diag minus stencils f64 2^20; C2 chain) and
the config-5 structure at 1/8 size as box calibration.  Prints ms per fwd+adj pair.  Env switches are read at jets_init."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, ctypes as C
import jets_b200 as B
B.init(0)
s = torch.cuda.current_stream()
B.check(B.lib.jets_stream_set(C.c_void_p(s.cuda_stream)))

def timed(fn, reps, warm):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s)
    for _ in range(reps): fn()
    b.record(s); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("JETS_B200_")) or "default"
res = {}
ONLY = os.environ.get("AB_ONLY", "c1,c4,c2,c5").split(",")
# C1, six rotating sets (>> L2)
def run_c1():
    n, NS = 1_000_000, 6
    sp = B.JetSpace(np.float64, n)
    sets = []
    for i in range(NS):
        W = B.rand(B.JetBSpace([sp] * 16), seed=1001 + 10 * i)
        A = B.blockop([[B.JopDiagonal(B.getblock(W, 1 + r + 4 * c)) for c in range(4)] for r in range(4)])
        sets.append((A, B.adjoint(A), B.rand(B.domain(A), seed=2 + i), B.zeros(B.range_(A)), B.zeros(B.domain(A)), W))
    cnt = [0]
    def step1():
        A_, At_, m_, d_, m2_, _ = sets[cnt[0] % NS]; cnt[0] += 1
        B.mul_(d_, A_, m_); B.mul_(m2_, At_, d_)
    res["c1_pair_us"] = round(timed(step1, 4 * NS * 5, 2 * NS) * 1e3, 2)
    del sets
# C4 operator
def run_c4():
    nb, n4 = 8, 1 << 20
    sp = B.JetSpace(np.float64, n4)
    W4 = [B.rand(B.JetBSpace([sp] * nb), seed=4001 + i) for i in range(4)]
    ops = []
    for W in W4:
        Bd = B.blockop([[B.JopDiagonal(B.getblock(W, i + 1)) if i == j else B.JopZeroBlock(sp, sp) for j in range(nb)] for i in range(nb)])
        Sd = B.blockop([[B.JopStencil(np.float64, n4, "lap") if i == j else B.JopZeroBlock(sp, sp) for j in range(nb)] for i in range(nb)])
        A4 = Bd - 0.5 * Sd
        ops.append((A4, B.adjoint(A4), B.rand(B.domain(A4), seed=5), B.zeros(B.range_(A4)), B.zeros(B.domain(A4))))
    c4 = [0]
    def step4():
        A_, At_, m_, d_, m2_ = ops[c4[0] % 4]; c4[0] += 1
        B.mul_(d_, A_, m_); B.mul_(m2_, At_, d_)
    res["c4op_pair_us"] = round(timed(step4, 100, 8) * 1e3, 2)
    del ops, W4
# C2 chain
def run_c2():
    n = 100_000_000
    T = np.float32
    sp = B.JetSpace(T, n)
    w, mo = B.rand(sp, seed=2001), B.rand(sp, seed=2002)
    G = B.JopDiagonal(w) @ B.JopStencil(T, n, "fdiff") @ B.JopPointwise(T, n, "square")
    Jc = B.jacobian(G, mo); Jt = B.adjoint(Jc)
    dm, dd, dm2 = B.rand(sp, seed=2003), B.zeros(sp), B.zeros(sp)
    def step2():
        B.mul_(dd, Jc, dm); B.mul_(dm2, Jt, dd)
    res["c2_pair_ms"] = round(timed(step2, 20, 3), 4)
    del G, Jc, Jt, w, mo, dm, dd, dm2
# C5 structure, 32 blocks (calibration of the box)
def run_c5():
    T = np.float32
    blk, nblk = 15_625_000, 32
    sp = B.JetSpace(T, blk)
    own = B.JetBSpace([sp] * nblk)
    W = B.rand(own, seed=1)
    Z = B.JopZeroBlock(sp, sp)
    A = B.blockop([[B.JopDiagonal(B.getblock(W, r + 1)) if r == c else B.JopStencil(T, blk, "fdiff") if c == r + 1 else
                    B.JopStencil(T, blk, "lap") if c == r - 1 else Z for c in range(nblk)] for r in range(nblk)])
    At = B.adjoint(A)
    x, d, m = B.rand(own, seed=2), B.zeros(own), B.zeros(own)
    def step5():
        B.mul_(d, A, x); B.mul_(m, At, d)
    res["c5s_pair_ms"] = round(timed(step5, 10, 3), 4)
for name, fn in (("c1", run_c1), ("c4", run_c4), ("c2", run_c2), ("c5", run_c5)):
    if name in ONLY:
        fn()
print(tag, res, flush=True)
