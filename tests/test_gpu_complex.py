"""GPU parity for the complex eltypes and symmetric spaces: ComplexF32/ComplexF64
JetSpace / JetBSpace vectors, the complex diagonal fixture (test/runtests.jl:3-8, :915-917), complex
operators under block / sum / composite / adjoint, and JetSSpace / SymmetricArray (:219-282).

Tolerances: 1e-12 relative (ComplexF64), 1e-5 (ComplexF32) against the numpy oracle.  Elementwise
paths are additionally bit-compared with the products written out in real arithmetic, one rounding
per operation and no FMA -- Julia's complex `*` (numpy's SIMD complex multiply may contract)."""
import math

import numpy as np
import pytest

from backends import OracleBackend, DeviceBackend

pytestmark = pytest.mark.gpu
TOL = {np.dtype(np.complex64): 1e-5, np.dtype(np.complex128): 1e-12}
CT = [np.complex64, np.complex128]


@pytest.fixture(scope="module")
def O():
    return OracleBackend()


@pytest.fixture(scope="module")
def D():
    return DeviceBackend()


def crand(g, n, T):
    rt = np.float32 if np.dtype(T) == np.complex64 else np.float64
    return (g.random(n).astype(rt) + 1j * g.random(n).astype(rt)).astype(T)


def cmul(a, b):
    """Julia's complex product in real arithmetic: (ar*br - ai*bi, ar*bi + ai*br)."""
    a, b = np.asarray(a), np.asarray(b)
    out = np.empty(np.broadcast(a, b).shape, dtype=np.result_type(a, b))
    out.real = a.real * b.real - a.imag * b.imag
    out.imag = a.real * b.imag + a.imag * b.real
    return out


def close(a, b, T):
    a, b = np.asarray(a).ravel(), np.asarray(b).ravel()
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) <= TOL[np.dtype(T)] * (nb if nb > 0 else 1.0)


def bits(a, b):
    a, b = np.atleast_1d(np.asarray(a)), np.atleast_1d(np.asarray(b))
    return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))


# ------------------------------------------------------------------ vectors -----------------
@pytest.mark.parametrize("T", CT)
def test_complex_block_arrays(O, D, T):  # runtests.jl:542-550 and :58-94 for complex T
    g = np.random.default_rng(11)
    B = D.B
    R = B.JetBSpace([B.JetSpace(T, 2), B.JetSpace(T, 2, 2), B.JetSpace(T, 2, 3), B.JetSpace(T, 2051)])
    n = len(R)
    xh, yh = crand(g, n, T), crand(g, n, T)
    x, y = B.to_device(xh, R), B.to_device(yh, R)
    assert x.dtype == np.dtype(T) and bits(x.to_host(), xh)
    assert [x.block_range(i) for i in (1, 2, 3, 4)] == [(1, 2), (3, 6), (7, 12), (13, n)]
    assert bits(B.getblock(x, 3).to_host().reshape(-1, order="F"), xh[6:12])
    # abs.(x): real eltype, same values as hypot
    ax = abs(x)
    assert ax.dtype == np.dtype(T).type(0).real.dtype and close(ax.to_host(), np.abs(xh), T)
    assert B.space(B.rand(R)) == R
    # dot conjugates the first argument (src/Jets.jl:853)
    d = B.dot(x, y)
    ref = np.vdot(xh.astype(np.complex128), yh.astype(np.complex128))
    assert abs(complex(d) - ref) <= TOL[np.dtype(T)] * abs(ref)
    assert abs(complex(B.dot(x, x)).imag) <= TOL[np.dtype(T)] * abs(complex(B.dot(x, x)).real)
    x64 = xh.astype(np.complex128)
    for p, want in ((2, np.linalg.norm(x64)), (1, np.sum(np.abs(x64))), (math.inf, np.max(np.abs(x64))),
                    (-math.inf, np.min(np.abs(x64))), (3.5, np.sum(np.abs(x64) ** 3.5) ** (1 / 3.5))):
        got = float(B.norm(x, p))
        assert abs(got - want) <= 10 * TOL[np.dtype(T)] * want, (p, got, want)
    z = B.to_device(np.where(np.arange(n) % 3 == 0, 0, xh).astype(T), R)
    assert float(B.norm(z, 0)) == np.count_nonzero(np.arange(n) % 3 != 0)
    with pytest.raises(B.JetsError):
        B.extrema(x)   # complex numbers are not ordered
    # broadcast: real coefficients scale both parts; complex coefficients are full products
    a, b_, c = 0.3, -1.7, 2.2
    rt = np.dtype(T).type(0).real.dtype.type
    lin = B.lincomb_(B.similar(x), [(a, x), (b_, y), (c, x)]).to_host()
    want = ((rt(a) * xh.view(rt) + rt(b_) * yh.view(rt)) + rt(c) * xh.view(rt)).view(T)
    assert bits(lin, want)
    ca, cb = 0.3 - 0.5j, 1.1 + 0.25j
    lin = B.lincomb_(B.similar(x), [(ca, x), (cb, y)]).to_host()
    want = cmul(np.dtype(T).type(ca), xh) + cmul(np.dtype(T).type(cb), yh)
    assert bits(lin, want.astype(T))
    assert bits((x * y).to_host(), cmul(xh, yh).astype(T))
    assert bits((x * (2 + 1j)).to_host(), cmul(np.dtype(T).type(2 + 1j), xh).astype(T))
    # fill! / setblock! with complex values; ones/zeros
    w = B.zeros(R)
    B.fill_(w, 3.0 - 2.0j)
    assert np.all(w.to_host() == np.dtype(T).type(3 - 2j))
    B.setblock_(w, 2, 1.5j)
    assert np.all(w.to_host()[2:6] == np.dtype(T).type(1.5j)) and w.to_host()[6] == np.dtype(T).type(3 - 2j)
    assert np.all(B.ones(R).to_host() == np.dtype(T).type(1))
    # scalar getindex / setindex! (1-based)
    x[5] = 7 - 1j
    assert x[5] == np.dtype(T).type(7 - 1j) and x.to_host()[4] == np.dtype(T).type(7 - 1j)


def test_complex_rand_is_partition_invariant(D):
    B = D.B
    n = 1000
    whole = B.rand(B.JetSpace(np.complex128, n), seed=5).to_host()
    assert np.all((whole.real >= 0) & (whole.real < 1) & (whole.imag >= 0) & (whole.imag < 1))
    assert abs(whole.real.mean() - 0.5) < 0.05 and abs(whole.imag.mean() - 0.5) < 0.05
    assert not np.array_equal(whole.real, whole.imag)
    zn = B.randn(B.JetSpace(np.complex128, 200000), seed=6).to_host()
    assert abs(np.mean(np.abs(zn) ** 2) - 1.0) < 0.02   # randn(ComplexF64) has unit variance in total


# ------------------------------------------------------------------ operators ---------------
@pytest.mark.parametrize("T", CT)
@pytest.mark.parametrize("n", [10, 4099])
def test_complex_diagonal_fixture(O, D, T, n):  # runtests.jl:3-8, :915-917
    g = np.random.default_rng(12)
    wh, mh, dh = crand(g, n, T), crand(g, n, T), crand(g, n, T)
    B, J = D.B, O.J
    A = B.JopDiagonal(wh)
    Ao = J.JopDiagonal(wh)
    m, d = B.to_device(mh), B.to_device(dh)
    f = (A * m).to_host()
    t = (A.T * d).to_host()
    assert bits(f, cmul(wh, mh)) and bits(t, cmul(np.conj(wh), dh))
    assert close(f, Ao * mh, T) and close(t, Ao.T * dh, T)
    lhs, rhs = B.dot_product_test(A, m, d)
    assert abs(complex(lhs) - complex(rhs)) <= 10 * TOL[np.dtype(T)] * abs(complex(lhs))
    assert B.plan_info(A)["engines"] == ["tma"]      # library-owned, aligned vectors: the TMA-staged interpreter
    l2, r2 = B.linearity_test(A)
    assert close(l2.to_host(), r2.to_host(), T)


@pytest.mark.parametrize("T", CT)
def test_complex_block_sum_composite(O, D, T):
    """3x2 block operator of complex diagonals, a complex scalar multiple, stencils and a zero block:
    forward / adjoint against the oracle, dot-product test, and convert(Array, A) against its
    conjugate transpose."""
    g = np.random.default_rng(13)
    n = 257

    def build(K, dev):
        w = [crand(np.random.default_rng(100 + i), n, T) for i in range(4)]
        Dg = [K.JopDiagonal(dev(v)) for v in w]
        S = K.JopStencil(T, n, "fdiff")
        Lp = K.JopStencil(T, n, "lap")
        a = (0.5 - 1.25j)
        rows = [[Dg[0], K.compose(Dg[1], S)],
                [a * Dg[2], K.JopZeroBlock(K.JetSpace(T, n), K.JetSpace(T, n))],
                [Lp - Dg[3], S]]
        return K.blockop(rows)

    B, J = D.B, O.J
    Ad = build(B, lambda v: B.to_device(v))
    Ao = build(J, lambda v: v)
    mh, dh = crand(g, 2 * n, T), crand(g, 3 * n, T)
    f = (Ad * B.to_device(mh, B.domain(Ad))).to_host()
    t = (Ad.T * B.to_device(dh, B.range_(Ad))).to_host()
    fo = J.to_array(Ao * J.reshape(mh.copy(), J.domain(Ao)))
    to = J.to_array(Ao.T * J.reshape(dh.copy(), J.range_(Ao)))
    assert close(f, fo, T) and close(t, to, T)
    lhs, rhs = B.dot_product_test(Ad, B.to_device(mh, B.domain(Ad)), B.to_device(dh, B.range_(Ad)))
    assert abs(complex(lhs) - complex(rhs)) <= 10 * TOL[np.dtype(T)] * abs(complex(lhs))
    if np.dtype(T) == np.complex128:
        M = B.to_matrix(Ad)
        Mt = B.to_matrix(Ad.T)
        assert np.linalg.norm(M.conj().T - Mt) <= 1e-12 * np.linalg.norm(M)
        assert np.linalg.norm(M @ mh - f) <= 1e-12 * np.linalg.norm(f)


@pytest.mark.parametrize("T", CT)
@pytest.mark.parametrize("n", [5000, 20_011])
def test_complex_tma_and_ldg_engines_agree_bitwise(D, T, n):
    """Complex spaces on both interpreter engines: operands staged in shared memory by cp.async.bulk (block starts
    16-byte aligned) or read with guarded loads (forced, or chosen by the planner for the odd ComplexF32 block length)
    -- stencil halos across the staged tile's pad, conjugating adjoint stages and row sums must give the same bits."""
    B = D.B
    g = np.random.default_rng(17)
    W = [[crand(g, n, T) for _ in range(3)] for _ in range(2)]
    m, d = crand(g, 3 * n, T), crand(g, 2 * n, T)
    a = 0.75 - 0.5j
    res = {}
    for eng in ("auto", "ldg"):
        B.set_fused_engine(eng)
        try:
            A = B.blockop([[B.JopDiagonal(W[r][c]) @ B.JopStencil(T, n, "lap") if c != 1 else a * (B.JopStencil(T, n, "fdiff") @ B.JopDiagonal(W[r][c]))
                            for c in range(3)] for r in range(2)])
            f = (A * B.to_device(m, B.domain(A))).to_host()
            t = (A.T * B.to_device(d, B.range_(A))).to_host()
            res[eng] = (f, t, B.plan_info(A)["engines"])
        finally:
            B.set_fused_engine("auto")
    aligned = (n * np.dtype(T).itemsize) % 16 == 0
    assert res["auto"][2] == (["tma"] if aligned else ["ldg"]) and res["ldg"][2] == ["ldg"]
    assert bits(res["auto"][0], res["ldg"][0]) and bits(res["auto"][1], res["ldg"][1])


def test_complex_pointwise_jacobian(O, D):
    """x -> x.^2 on a complex space: J = 2 mo .* dm, J' = conj(2 mo) .* d."""
    g = np.random.default_rng(14)
    T, n = np.complex128, 513
    B = D.B
    F = B.JopPointwise(T, n, "square")
    mo, dm, d = crand(g, n, T), crand(g, n, T), crand(g, n, T)
    assert close((F * B.to_device(mo)).to_host(), mo * mo, T)
    Jc = B.jacobian(F, B.to_device(mo))
    assert close((Jc * B.to_device(dm)).to_host(), 2 * mo * dm, T)
    assert close((Jc.T * B.to_device(d)).to_host(), np.conj(2 * mo) * d, T)
    lhs, rhs = B.dot_product_test(Jc, B.to_device(dm), B.to_device(d))
    assert abs(complex(lhs) - complex(rhs)) <= 1e-11 * abs(complex(lhs))
    with pytest.raises(B.JetsError):
        B.JopPointwise(T, n, "exp")


# ------------------------------------------------------------------ symmetric spaces --------
def indexmap(I):  # runtests.jl:219-225
    return tuple(I) if I[0] < 5 else (I[0] - 4, I[1])


def test_symmetric_space(O, D):  # runtests.jl:227-258
    B, J = D.B, O.J
    R = B.JetSSpace(np.complex128, (8, 4), (4, 4), indexmap)
    assert R.size() == (8, 4) and R.eltype == np.complex128
    assert np.array_equal(B.ones(R).to_host(), np.ones((8, 4), dtype=np.complex128))
    assert np.array_equal(B.zeros(R).to_host(), np.zeros((8, 4), dtype=np.complex128))
    assert B.rand(R).shape == (8, 4) and B.Array(R).shape == (8, 4) and B.Array(R).dtype == np.complex128
    x = B.rand(R, seed=21)
    z = B.similar(x)
    assert z.issymmetric and z.space == R
    y = x.A                                     # parent, size M
    assert y.shape == (4, 4)
    xo = J.SymmetricArray(np.asfortranarray(y.copy()), (8, 4), indexmap)
    assert np.array_equal(x.to_host(), xo.full())
    for p in (2, 1, math.inf, 3.5):
        assert np.isclose(float(B.norm(x, p)), J.norm(xo, p), rtol=1e-12), p
    assert np.isclose(float(B.norm(x)), math.sqrt(2 * np.linalg.norm(y) ** 2), rtol=1e-12)
    assert np.isclose(float(B.norm(x, 1)), 2 * np.sum(np.abs(y)), rtol=1e-12)
    assert np.isclose(float(B.norm(x, math.inf)), np.max(np.abs(y)), rtol=1e-12)
    x[(1, 1)] = 0
    x[(6, 1)] = 0
    assert float(B.norm(x, 0)) == 2 * np.count_nonzero(x.A)
    assert B.space(B.rand(R)) == R
    for i in range(1, 33):
        x[i] = i + 1j * i
        assert x[i] == i + 1j * i
    x[(7, 2)] = 3 + 4j
    assert x.A[2, 1] == 3 - 4j and x[(7, 2)] == 3 + 4j and x[(3, 2)] == 3 - 4j


@pytest.mark.parametrize("T", [np.float32, np.float64])
def test_symmetric_space_real_eltype(O, D, T):
    """JetSSpace with a real eltype (src/Jets.jl:408-441: T is "usually" complex, not necessarily): conj is the
    identity, the norms still count every stored element once per logical position it stands for."""
    B, J = D.B, O.J
    R = B.JetSSpace(T, (8, 4), (4, 4), indexmap)
    x = B.rand(R, seed=22)
    y = x.A
    assert y.shape == (4, 4) and y.dtype == np.dtype(T) and x.to_host().shape == (8, 4)
    xo = J.SymmetricArray(np.asfortranarray(y.copy()), (8, 4), indexmap)
    assert np.array_equal(x.to_host(), xo.full())
    tol = 1e-6 if T == np.float32 else 1e-12
    for p in (2, 1, math.inf, -math.inf, 3.5):
        assert np.isclose(float(B.norm(x, p)), float(J.norm(xo, p)), rtol=tol), p
    assert np.isclose(float(B.norm(x)), math.sqrt(2 * np.linalg.norm(y.astype(np.float64)) ** 2), rtol=tol)
    x[(1, 1)] = 0
    assert float(B.norm(x, 0)) == 2 * np.count_nonzero(x.A)
    x[(7, 2)] = 3
    assert x.A[2, 1] == 3 and x[(3, 2)] == 3
    z = x * 0.5 + B.rand(R, seed=23) * 0.25
    assert z.issymmetric and z.space == R


def test_symmetric_space_broadcast(D):  # runtests.jl:260-282
    B = D.B
    R = B.JetSSpace(np.complex128, (8, 4), (4, 4), indexmap)
    u, v, w = B.rand(R, seed=1), B.rand(R, seed=2), B.rand(R, seed=3)
    a, b, c = 0.25, 0.5, 0.125
    x = u * a + v * b + w * c
    assert x.issymmetric and x.space == R
    assert np.allclose(x.A, a * u.A + b * v.A + c * w.A, rtol=1e-15, atol=0)
    y = B.zeros(R).assign(x)
    assert y.issymmetric and np.array_equal(y.A, x.A)


# ------------------------------------------------------------------ leaves + state (rank 4) --
@pytest.mark.parametrize("T", [np.float32, np.float64, np.complex128])
def test_restriction(O, D, T):
    """d = m[idx], adjoint scatter; alone, inside a composite with fused stages on both sides, and as
    a block of a JopBlock -- bit-exact against the oracle (pure data movement + elementwise)."""
    g = np.random.default_rng(31)
    n, k = 5003, 1777
    idx = g.permutation(n)[:k] + 1
    rt = np.dtype(T).type(0).real.dtype
    mk = (lambda s: crand(g, s, T)) if np.dtype(T).kind == "c" else (lambda s: g.random(s).astype(T))
    w1, w2, mh, dh = mk(n), mk(k), mk(n), mk(k)

    def build(K, dev):
        R = K.JopRestriction(T, n, idx)
        chain = K.compose(K.JopDiagonal(dev(w2)), K.compose(R, K.JopDiagonal(dev(w1))))
        blk = K.blockop([[R, K.JopDiagonal(dev(w2))], [K.JopStencil(T, n, "fdiff"), K.JopZeroBlock(K.JetSpace(T, k), K.JetSpace(T, n))]])
        return R, chain, blk

    B, J = D.B, O.J
    Rd, Cd, Bd = build(B, lambda v: B.to_device(v))
    Ro, Co, Bo = build(J, lambda v: v)
    assert bits((Rd * B.to_device(mh)).to_host(), Ro * mh)
    assert bits((Rd.T * B.to_device(dh)).to_host(), Ro.T * dh)
    assert "gather" in B.plan_info(Rd)["engines"]
    if np.dtype(T).kind != "c":       # complex products: numpy's SIMD multiply may contract
        assert bits((Cd * B.to_device(mh)).to_host(), Co * mh)
        assert bits((Cd.T * B.to_device(dh)).to_host(), Co.T * dh)
    else:
        assert close((Cd * B.to_device(mh)).to_host(), Co * mh, T) and close((Cd.T * B.to_device(dh)).to_host(), Co.T * dh, T)
    xin = np.concatenate([mh, dh])
    yin = np.concatenate([dh, mh])
    f = (Bd * B.to_device(xin, B.domain(Bd))).to_host()
    t = (Bd.T * B.to_device(yin, B.range_(Bd))).to_host()
    fo = J.to_array(Bo * J.reshape(xin.copy(), J.domain(Bo)))
    to = J.to_array(Bo.T * J.reshape(yin.copy(), J.range_(Bo)))
    tol = 1e-5 if rt == np.float32 else 1e-12
    assert np.linalg.norm(f - fo) <= tol * np.linalg.norm(fo) and np.linalg.norm(t - to) <= tol * np.linalg.norm(to)
    lhs, rhs = B.dot_product_test(Bd, B.to_device(xin, B.domain(Bd)), B.to_device(yin, B.range_(Bd)))
    assert abs(complex(lhs) - complex(rhs)) <= 10 * tol * abs(complex(lhs))
    with pytest.raises(B.JetsError):
        B.JopRestriction(T, 10, [1, 2, 2])      # duplicate index
    with pytest.raises(B.JetsError):
        B.JopRestriction(T, 10, [1, 11])        # outside the domain


@pytest.mark.parametrize("fn", ["log", "atan"])
def test_new_pointwise_functions(O, D, fn):
    g = np.random.default_rng(32)
    n = 3001
    B, J = D.B, O.J
    mo, dm = g.random(n) + 0.5, g.random(n)
    Fd, Fo = B.JopPointwise(np.float64, n, fn), J.JopPointwise(np.float64, n, fn)
    assert np.allclose((Fd * B.to_device(mo)).to_host(), Fo * mo, rtol=1e-12, atol=0)
    Jd, Jo = B.jacobian(Fd, B.to_device(mo)), J.jacobian(Fo, mo)
    assert np.allclose((Jd * B.to_device(dm)).to_host(), Jo * dm, rtol=1e-12, atol=0)
    assert np.allclose((Jd.T * B.to_device(dm)).to_host(), Jo.T * dm, rtol=1e-12, atol=0)
    muobs, muexp = B.linearization_test(Fd, B.to_device(mo))
    assert np.isclose(muobs[-1], muexp[-1], rtol=0.1)


def test_state_plumbing(D):  # state/state!/perfstat/close: src/Jets.jl:264-290, :591-623
    B = D.B
    g = np.random.default_rng(33)
    n = 1000
    w = g.random(n)
    A = B.JopDiagonal(w)
    S = B.JopScale(np.float64, n, 2.5)
    Cmp = B.compose(S, A)
    assert np.array_equal(B.state(A, "diagonal").to_host(), w) and B.state(S, "a") == 2.5
    assert np.array_equal(B.state(Cmp, "diagonal").to_host(), w) and B.state(Cmp, "a") == 2.5   # :607-623
    with pytest.raises(KeyError):
        B.state(Cmp, "nope")
    with pytest.raises(KeyError):
        B.state(B.compose(A, B.JopDiagonal(w)), "diagonal")    # ambiguous
    m = g.random(n)
    assert np.array_equal(Cmp * m, 2.5 * (w * m))
    w2 = g.random(n)
    B.state_(A, {"diagonal": w2})              # state!: the kernels read the new values
    assert np.array_equal(Cmp * m, 2.5 * (w2 * m))
    with pytest.raises(B.JetsError):           # constants compiled into the device operator cannot be swapped silently
        B.state_(S, {"a": 3.0})
    B.state_(S, {"note": "user metadata"})     # anything else is host-side state, as in the reference
    assert B.state(S, "note") == "user metadata" and B.state(S, "a") == 2.5
    ps = B.perfstat(Cmp)
    assert ps["engines"] == ["tma"] and ps["launches"] == 1
    assert B.close(A) is False and B.close(B.compose(S, B.JopDiagonal(w))) is None


# ------------------------------------------------------------------ complex dense blocks ----
@pytest.mark.parametrize("T", CT)
@pytest.mark.parametrize("shape", [(1, 1, 300, 200), (2, 3, 128, 64), (3, 2, 77, 131), (1, 2, 1031, 516), (2, 1, 64, 2500)])
def test_complex_dense_block_gemv(O, D, T, shape):
    """Complex matrices as operators: d = A m, m = A' d with A' the conjugate transpose (_matmul_df!/_df'!
    src/Jets.jl:573-574) under a block operator, ragged shapes included."""
    nr, nc, rows, cols = shape
    g = np.random.default_rng(41)
    Bm = [[(crand(g, rows * cols, T) - T(0.5 + 0.25j)).reshape(rows, cols) for _ in range(nc)] for _ in range(nr)]
    m = crand(g, nc * cols, T)
    d = crand(g, nr * rows, T)

    def scn(K):
        A = K.blockop([[K.JopDense(Bm[r][c]) for c in range(nc)] for r in range(nr)])
        lhs, rhs = K.dot_product_test(A, K.arr(m, K.domain(A)), K.arr(d, K.range_(A)))
        return {"f": K.host(A * K.arr(m, K.domain(A))), "t": K.host(K.adjoint(A) * K.arr(d, K.range_(A))),
                "dpt": np.array([lhs, rhs], dtype=np.complex128)}
    o, dv = scn(O), scn(D)
    M = np.block([[b.astype(np.complex128) for b in row] for row in Bm])
    assert close(dv["f"], M @ m.astype(np.complex128), T) and close(dv["f"], o["f"], T)
    assert close(dv["t"], M.conj().T @ d.astype(np.complex128), T) and close(dv["t"], o["t"], T)
    assert abs(dv["dpt"][0] - dv["dpt"][1]) <= TOL[np.dtype(T)] * abs(dv["dpt"][0] + dv["dpt"][1])
    Ad = D.blockop([[D.JopDense(Bm[r][c]) for c in range(nc)] for r in range(nr)])
    Ad * D.arr(m, D.domain(Ad))
    info = D.B.plan_info(Ad)
    assert info["engines"] == ["gemv"] and info["launches"] == 1


@pytest.mark.parametrize("T", CT)
def test_complex_dense_in_composition_and_multi_rhs(O, D, T):
    """A complex matrix composed with a complex diagonal (staged through a device temporary where the reference
    allocates one, :525-539), and a matrix of right-hand sides (one GEMV per column: tcgen05 is Float32-only)."""
    g = np.random.default_rng(42)
    B, J = D.B, O.J
    rows, cols, nrhs = 150, 90, 3
    A = (crand(g, rows * cols, T) - T(0.5)).reshape(rows, cols)
    w = crand(g, rows, T)
    m, d = crand(g, cols, T), crand(g, rows, T)
    Cd = B.JopDiagonal(w) @ B.JopDense(A)
    Co = J.JopDiagonal(w) @ J.JopDense(A)
    assert close((Cd * B.to_device(m)).to_host(), Co * m, T)
    assert close((Cd.T * B.to_device(d)).to_host(), Co.T * d, T)
    X = crand(g, cols * nrhs, T).reshape((cols, nrhs), order="F")
    op = B.JopDense(A, nrhs=nrhs)
    Y = op * B.to_device(X, B.domain(op))
    assert Y.shape == (rows, nrhs) and close(Y.to_host(), A.astype(np.complex128) @ X.astype(np.complex128), T)
    Xb = op.T * Y
    assert close(Xb.to_host(), A.astype(np.complex128).conj().T @ Y.to_host().astype(np.complex128), T)
    assert B.plan_info(op)["engines"] == ["gemv"]
    # accumulate on top of a fused term inside one block row (Q1 order: fused SETs, dense adds)
    Ad = B.blockop([[B.JopDense(A), B.JopDiagonal(w)]])
    Ao = J.blockop([[J.JopDense(A), J.JopDiagonal(w)]])
    x = np.concatenate([m, d])
    assert close((Ad * B.to_device(x, B.domain(Ad))).to_host(), J.to_array(Ao * J.reshape(x.copy(), J.domain(Ao))), T)
    assert close((Ad.T * B.to_device(d)).to_host(), J.to_array(Ao.T * d), T)
    # signed sums of complex matrices: the second GEMV subtracts from what the first one stored
    A2 = (crand(g, rows * cols, T) - T(0.5j)).reshape(rows, cols)
    Sd, So = B.JopDense(A) - B.JopDense(A2), J.JopDense(A) - J.JopDense(A2)
    assert close((Sd * B.to_device(m)).to_host(), So * m, T) and close((Sd.T * B.to_device(d)).to_host(), So.T * d, T)
