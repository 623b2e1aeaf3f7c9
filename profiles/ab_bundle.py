"""A/B timing of the fused block-apply engines on the BASELINE configs (kernel-only, CUDA events).
usage: python profiles/ab_bundle.py [c1] [c2] [c4] [c5] [c5s]   (engine options come from JETS_B200_* env)
Prints one JSON line per (config, engine)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import jets_b200 as B  # noqa: E402
import ctypes as C  # noqa: E402

B.init(0)
stream = torch.cuda.current_stream()
B.check(B.lib.jets_stream_set(C.c_void_p(stream.cuda_stream)))
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")
PEAK = 6650.0
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timeit(fn, steps, warmup, do_flush):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(steps):
        if do_flush:
            flush.fill_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        fn()
        b.record(stream)
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def build(name):
    if name == "c1":
        n, T = 1_000_000, np.float64
        sp = B.JetSpace(T, n)
        W = B.rand(B.JetBSpace([sp] * 16), seed=1001)
        A = B.blockop([[B.JopDiagonal(B.getblock(W, 1 + r + 4 * c)) for c in range(4)] for r in range(4)])
        return A, 192e6, True, W
    if name == "c2":
        n, T = 100_000_000, np.float32
        sp = B.JetSpace(T, n)
        w, mo = B.rand(sp, seed=2001), B.rand(sp, seed=2002)
        G = B.JopDiagonal(w) @ B.JopStencil(T, n, "fdiff") @ B.JopPointwise(T, n, "square")
        return B.jacobian(G, mo), 1.6e9, False, (w, mo)
    if name == "c4":
        nb, n, T = 8, 1 << 20, np.float64
        sp = B.JetSpace(T, n)
        W = B.rand(B.JetBSpace([sp] * nb), seed=4001)
        Bd = B.blockop([[B.JopDiagonal(B.getblock(W, i + 1)) if i == j else B.JopZeroBlock(sp, sp) for j in range(nb)] for i in range(nb)])
        Sd = B.blockop([[B.JopStencil(T, n, "lap") if i == j else B.JopZeroBlock(sp, sp) for j in range(nb)] for i in range(nb)])
        return Bd - 0.5 * Sd, 3 * nb * n * 8, True, W
    if name in ("c5", "c5s"):
        nb, blk, T = (256, 15_625_000, np.float32) if name == "c5" else (32, 15_625_000, np.float32)
        sp = B.JetSpace(T, blk)
        W = B.zeros(B.JetBSpace([sp] * nb))
        B.check(B.lib.jets_buf_rand(W._h, 5001, 0, 0))
        Z = B.JopZeroBlock(sp, sp)
        Sup, Slo = B.JopStencil(T, blk, "fdiff"), B.JopStencil(T, blk, "lap")
        A = B.blockop([[B.JopDiagonal(B.getblock(W, r + 1)) if r == c else Sup if c == r + 1 else Slo if c == r - 1 else Z
                        for c in range(nb)] for r in range(nb)])
        return A, 3.0 * nb * blk * 4, False, W
    raise SystemExit(f"unknown config {name}")


ROT = int(os.environ.get("AB_ROTATE", "0"))   # >0: cycle over ROT independent operator/vector sets instead of flushing L2
for name in (sys.argv[1:] or ["c1", "c2", "c4", "c5s"]):
    for eng in ("auto", "tma_nocache"):
        B.set_fused_engine(eng)
        A, nbytes, do_flush, keep = build(name)
        At = B.adjoint(A)
        m, d = B.rand(B.domain(A), seed=7), B.zeros(B.range_(A))
        m2 = B.zeros(B.domain(A))
        steps = 30 if nbytes < 4e9 else 8
        if ROT > 0 and do_flush:
            sets = [(A, At, m, d, m2, keep)]
            for i in range(1, ROT):
                Ai, _, _, ki = build(name)
                sets.append((Ai, B.adjoint(Ai), B.rand(B.domain(A), seed=70 + i), B.zeros(B.range_(A)), B.zeros(B.domain(A)), ki))
            cnt = [0, 0]

            def fwd():
                s_ = sets[cnt[0] % ROT]; cnt[0] += 1
                B.mul_(s_[3], s_[0], s_[2])

            def adj():
                s_ = sets[cnt[1] % ROT]; cnt[1] += 1
                B.mul_(s_[4], s_[1], s_[3])
            f_med, f_min = timeit(fwd, steps, 2 * ROT, False)
            t_med, t_min = timeit(adj, steps, 2 * ROT, False)
            cnt[0] = cnt[1] = 0
            p_med, p_min = timeit(lambda: (fwd(), adj(), fwd(), adj()), steps, ROT, False)   # 4 dependent launches, no events between
        else:
            f_med, f_min = timeit(lambda: B.mul_(d, A, m), steps, 3, do_flush)
            t_med, t_min = timeit(lambda: B.mul_(m2, At, d), steps, 3, do_flush)
            p_med, p_min = timeit(lambda: (B.mul_(d, A, m), B.mul_(m2, At, d), B.mul_(d, A, m), B.mul_(m2, At, d)), steps, 3, do_flush)
        lhs, rhs = B.dot_product_test(A, m, B.rand(B.range_(A), seed=8))
        print(json.dumps({"config": name, "engine": eng, "info": B.plan_info(A),
                          "fwd_ms": round(f_med, 4), "adj_ms": round(t_med, 4), "fwd_min_ms": round(f_min, 4), "per_apply_in_4chain_ms": round(p_med / 4, 4), "chain_frac": round(nbytes / (p_med / 4) / 1e6 / PEAK, 4),
                          "fwd_gbs": round(nbytes / f_med / 1e6, 1), "adj_gbs": round(nbytes / t_med / 1e6, 1),
                          "frac": round(nbytes / ((f_med + t_med) / 2) / 1e6 / PEAK, 4),
                          "dpt": float(abs(lhs - rhs) / abs(lhs + rhs)),
                          "env": {k: v for k, v in os.environ.items() if k.startswith("JETS_B200")}}), flush=True)
        del A, At, m, d, m2, keep
B.set_fused_engine("auto")
