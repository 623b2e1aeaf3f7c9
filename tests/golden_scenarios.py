"""Scenario definitions shared by tests/golden/make_golden.py (writes the oracle's outputs) and
tests/test_golden.py (replays them through the oracle on CPU and through the C ABI on the GPU).  Every
scenario is a pure function of a backend K (tests/backends.py) and fixed seeds; it returns a dict of
arrays.  Keys ending in "_red" are reductions / dense products (tolerance-compared); all others are
elementwise / stencil / data-movement results (bit-compared)."""
import math

import numpy as np


def _dev(K, v, T=None):
    v = np.asarray(v) if T is None else np.asarray(v, dtype=T)
    return K.arr(v, K.JetSpace(v.dtype, v.size)) if K.name == "device" else v


def _apply(K, A, m, d):
    R, Dm = K.range_(A), K.domain(A)
    f = K.host(K.mul_(K.zeros(R), A, K.arr(m, Dm)))
    t = K.host(K.mul_(K.zeros(Dm), K.adjoint(A), K.arr(d, R)))
    return np.asarray(f).reshape(-1, order="F"), np.asarray(t).reshape(-1, order="F")


def scn_block_diag(K, T):  # config 1 in miniature: 4x4 diagonal JopLn blocks (runtests.jl:622-666)
    g = np.random.default_rng(101)
    n = 257
    W = [[g.random(n).astype(T) for _ in range(4)] for _ in range(4)]
    A = K.blockop([[K.JopDiagonal(_dev(K, W[r][c])) for c in range(4)] for r in range(4)])
    m, d = g.random(4 * n).astype(T), g.random(4 * n).astype(T)
    f, t = _apply(K, A, m, d)
    return {"fwd": f, "adj": t}


def scn_chain(K, T):  # config 2 in miniature: diagonal ∘ fdiff ∘ jacobian(x^2)
    g = np.random.default_rng(102)
    n = 1031
    w, mo = g.random(n).astype(T), g.random(n).astype(T)
    F = K.compose(K.JopDiagonal(_dev(K, w)), K.compose(K.JopStencil(T, n, "fdiff"), K.JopPointwise(T, n, "square")))
    Jc = K.jacobian(F, _dev(K, mo))
    m, d = g.random(n).astype(T), g.random(n).astype(T)
    f, t = _apply(K, Jc, m, d)
    nl = np.asarray(K.host(K.mul_(K.zeros(K.range_(F)), F, K.arr(m, K.domain(F))))).reshape(-1)
    return {"fwd": f, "adj": t, "nonlinear": nl}


def scn_tridiag(K, T):  # config 5 in miniature: block-tridiagonal, diagonal + stencils + zero blocks
    g = np.random.default_rng(103)
    nb, n = 6, 515
    W = [g.random(n).astype(T) for _ in range(nb)]
    sp = K.JetSpace(T, n)

    def blk(r, c):
        if r == c:
            return K.JopDiagonal(_dev(K, W[r]))
        if c == r + 1:
            return K.JopStencil(T, n, "fdiff")
        if c == r - 1:
            return K.JopStencil(T, n, "lap")
        return K.JopZeroBlock(sp, sp)
    A = K.blockop([[blk(r, c) for c in range(nb)] for r in range(nb)])
    m, d = g.random(nb * n).astype(T), g.random(nb * n).astype(T)
    f, t = _apply(K, A, m, d)
    return {"fwd": f, "adj": t}


def scn_sum_scaled(K, T):  # config 4's operator in miniature: B - 0.5*S
    g = np.random.default_rng(104)
    nb, n = 3, 401
    W = [g.random(n).astype(T) for _ in range(nb)]
    sp = K.JetSpace(T, n)
    Z = lambda: K.JopZeroBlock(sp, sp)
    Bd = K.blockop([[K.JopDiagonal(_dev(K, W[i])) if i == j else Z() for j in range(nb)] for i in range(nb)])
    Sd = K.blockop([[K.JopStencil(T, n, "lap") if i == j else Z() for j in range(nb)] for i in range(nb)])
    A = Bd - 0.5 * Sd
    m, d = g.random(nb * n).astype(T), g.random(nb * n).astype(T)
    f, t = _apply(K, A, m, d)
    return {"fwd": f, "adj": t}


def scn_vectors(K, T):  # BlockArray reductions and broadcasts (runtests.jl:512-600)
    g = np.random.default_rng(105)
    R = K.JetBSpace([K.JetSpace(T, 2), K.JetSpace(T, 2, 2), K.JetSpace(T, 2, 3), K.JetSpace(T, 1031)])
    n = 2 + 4 + 6 + 1031
    xh, yh = g.standard_normal(n).astype(T), g.standard_normal(n).astype(T)
    x, y = K.arr(xh, R), K.arr(yh, R)
    out = {"lin": K.host(K.lincomb([(0.25, x), (-1.5, y), (3.0, x)])), "had": K.host(K.hadamard(x, y))}
    out["norms_red"] = np.array([float(K.norm(x, p)) for p in (2, 1, 0, math.inf, -math.inf, 3.5)])
    out["dot_red"] = np.array([float(K.dot(x, y))])
    mn, mx = K.extrema(x)
    out["extrema"] = np.array([mn, mx], dtype=T)
    return out


def scn_dense(K, T):  # matrix blocks (fixture JopBaz, runtests.jl:27-33) in a 2x3 JopBlock
    g = np.random.default_rng(106)
    shapes = [(130, 70), (130, 33), (130, 257)], [(65, 70), (65, 33), (65, 257)]
    mats = [[g.random(s).astype(T) for s in row] for row in shapes]
    A = K.blockop([[K.JopDense(M) for M in row] for row in mats])
    m, d = g.random(70 + 33 + 257).astype(T), g.random(130 + 65).astype(T)
    f, t = _apply(K, A, m, d)
    return {"fwd_red": f, "adj_red": t}


def scn_restriction(K, T):
    g = np.random.default_rng(107)
    n = 1001
    idx = g.permutation(n)[:400] + 1
    w = g.random(400).astype(T)
    A = K.compose(K.JopDiagonal(_dev(K, w)), K.JopRestriction(T, n, idx))
    m, d = g.random(n).astype(T), g.random(400).astype(T)
    f, t = _apply(K, A, m, d)
    return {"fwd": f, "adj": t}


def _crand(g, n, T):
    rt = np.float32 if np.dtype(T) == np.complex64 else np.float64
    return ((g.random(n).astype(rt) - rt(0.5)) + 1j * (g.random(n).astype(rt) - rt(0.5))).astype(T)


def scn_cdense(K, T):  # complex matrices as operators: d = A m, m = A' d with the conjugate transpose (src/Jets.jl:573-574)
    g = np.random.default_rng(108)
    shapes = [(77, 40), (77, 131)], [(130, 40), (130, 131)]
    mats = [[_crand(g, s[0] * s[1], T).reshape(s) for s in row] for row in shapes]
    A = K.blockop([[K.JopDense(M) for M in row] for row in mats])
    m, d = _crand(g, 40 + 131, T), _crand(g, 77 + 130, T)
    f, t = _apply(K, A, m, d)
    return {"fwd_red": f, "adj_red": t}


def scn_cblock(K, T):  # complex diagonals, a complex scalar multiple and stencils in a 2x2 JopBlock (conj in the adjoint)
    g = np.random.default_rng(109)
    n = 512
    w = [_crand(g, n, T) for _ in range(3)]
    A = K.blockop([[K.JopDiagonal(_dev(K, w[0])), K.compose(K.JopDiagonal(_dev(K, w[1])), K.JopStencil(T, n, "fdiff"))],
                   [(0.5 - 1.25j) * K.JopDiagonal(_dev(K, w[2])), K.JopStencil(T, n, "lap")]])
    m, d = _crand(g, 2 * n, T), _crand(g, 2 * n, T)
    f, t = _apply(K, A, m, d)
    # numpy's SIMD complex multiply may contract where the device rounds every real operation: tolerance keys
    return {"fwd_red": f, "adj_red": t}


SCENARIOS = {
    "block_diag_f64": (scn_block_diag, np.float64), "block_diag_f32": (scn_block_diag, np.float32),
    "chain_f32": (scn_chain, np.float32), "chain_f64": (scn_chain, np.float64),
    "tridiag_f32": (scn_tridiag, np.float32), "tridiag_f64": (scn_tridiag, np.float64),
    "sum_scaled_f64": (scn_sum_scaled, np.float64),
    "vectors_f32": (scn_vectors, np.float32), "vectors_f64": (scn_vectors, np.float64),
    "dense_f32": (scn_dense, np.float32), "dense_f64": (scn_dense, np.float64),
    "restriction_f64": (scn_restriction, np.float64),
    "cdense_c64": (scn_cdense, np.complex64), "cdense_c128": (scn_cdense, np.complex128),
    "cblock_c64": (scn_cblock, np.complex64), "cblock_c128": (scn_cblock, np.complex128),
}


def run_scenario(K, name):
    fn, T = SCENARIOS[name]
    return fn(K, T)
