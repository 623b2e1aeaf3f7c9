"""Pins the CPU oracle against every known-answer identity the reference's own test-suite holds
for the hot path (/root/reference/test/runtests.jl, line ranges cited per test).  The reference
stores no golden vectors and uses unseeded ``rand``; each testset is restated on seeded inputs.
``≈`` in the reference is Julia's norm-wise isapprox with rtol=sqrt(eps); we assert tighter.
"""
import math
import os

import numpy as np
import pytest

from oracle import jets_oracle as J
from oracle.jets_oracle import (JetSpace, JetBSpace, JopLn, JopNl, JopAdjoint, adjoint, compose,
                                jacobian, jacobian_, mul_, state, domain, range_)

RT = 1e-13


def approx(a, b, rtol=RT):
    a = J.to_array(a) if isinstance(a, J.BlockArray) else np.asarray(a)
    b = J.to_array(b) if isinstance(b, J.BlockArray) else np.asarray(b)
    na, nb = np.linalg.norm(a.ravel()), np.linalg.norm(b.ravel())
    return np.linalg.norm((a - b).ravel()) <= rtol * max(na, nb)


# ---- fixtures: test/runtests.jl:3-56 -------------------------------------------------------
def JopFoo(diag):  # :3-8
    def df(d, m, *, diagonal, **kw):
        d[...] = diagonal * m
        return d

    def dft(m, d, *, diagonal, **kw):
        m[...] = np.conj(diagonal) * d
        return m
    spc = JetSpace(diag.dtype, diag.size)
    return JopLn(df=df, dft=dft, dom=spc, rng=spc, s={"diagonal": diag})


def JopBar(n):  # :20-25
    def f(d, m, **kw):
        d[...] = m ** 2
        return d

    def df(dd, dm, *, mo, **kw):
        dd[...] = 2 * mo * dm
        return dd
    spc = JetSpace(np.float64, n)
    return JopNl(f=f, df=df, dom=spc, rng=spc)


def JopBaz(A):  # :27-33
    def df(d, m, *, A, **kw):
        d[...] = A @ m
        return d

    def dft(m, d, *, A, **kw):
        m[...] = A.conj().T @ d
        return m
    return JopLn(df=df, dft=dft, dom=JetSpace(A.dtype, A.shape[1]),
                 rng=JetSpace(A.dtype, A.shape[0]), s={"A": A})


def JopFooBar(n, g):  # :35-39
    def df(d, m, *, A, **kw):
        d[...] = A * m
        return d
    spc = JetSpace(np.float64, n, n)
    return JopLn(df=df, dom=spc, rng=spc, s={"A": g.random((n, n))})


def JopRosenbrock():  # :41-50
    def f(d, m, **kw):
        d[...] = [1 - m[0], 10 * (m[1] - m[0] ** 2)]
        return d

    def df(d, m, *, J, **kw):
        d[...] = J @ m
        return d

    def dft(m, d, *, J, **kw):
        m[...] = J.T @ d
        return m

    def up(m, s):
        s["J"][1, 0] = -20.0 * m[0]
    spc = JetSpace(np.float64, 2)
    return JopNl(f=f, df=df, dft=dft, upstate=up, dom=spc, rng=spc,
                 s={"J": np.array([[-1.0, 0.0], [0.0, 10.0]])})


@pytest.fixture
def g():
    return np.random.default_rng(20261017)


# ---- spaces: runtests.jl:58-94 ------------------------------------------------------------
@pytest.mark.parametrize("n", [(2,), (2, 3), (2, 3, 4)])
@pytest.mark.parametrize("T", [np.float32, np.float64, np.complex64, np.complex128])
def test_jetspace(n, T, g):
    R = JetSpace(T, *n)
    assert R.size() == n and R.eltype == np.dtype(T) and R.ndims == len(n)
    x = J.rand(R, g)
    assert x.dtype == np.dtype(T) and x.shape == n
    assert np.all(J.ones(R) == 1) and np.all(J.zeros(R) == 0)
    assert J.Array(R).shape == n
    assert R == J.space(J.rand(R, g))
    assert R.similar((0,) * len(n)) == JetSpace(T, *((0,) * len(n)))


def test_jet_construction(g):  # :96-124
    def f(d, m, *, a, **kw):
        d[...] = (a * m ** 2).reshape(d.shape, order="F")
        return d

    def df(dd, dm, *, a, mo, **kw):
        dd[...] = (2 * a * mo * dm).reshape(dd.shape, order="F")
        return dd
    a = g.random(20)
    jt = J.Jet(dom=JetSpace(np.float64, 20), rng=JetSpace(np.float64, 10, 2), f=f, df=df, dft=df,
               s={"a": a})
    assert domain(jt) == JetSpace(np.float64, 20) and range_(jt) == JetSpace(np.float64, 10, 2)
    assert J.point(jt).shape == (0,)
    mo = J.rand(domain(jt), g)
    J.point_(jt, mo)
    assert J.point(jt) is mo
    assert J.shape(jt) == ((10, 2), (20,)) and J.size(jt) == (20, 20)
    assert J.close(jt) is False
    with pytest.raises(ValueError):
        J.Jet(dom=JetSpace(np.float64, 2), rng=JetSpace(np.float64, 2))


def test_linear_operator(g):  # :126-168
    diag = g.random(10)
    A = JopFoo(diag)
    m = g.random(10)
    d = A * m
    assert np.array_equal(d, diag * m)
    a = A.T * d
    assert np.array_equal(a, diag * d)
    d[...] = 0
    mul_(d, A, m)
    assert np.array_equal(d, diag * m)
    assert J.size(A) == (10, 10) and J.shape(A) == ((10,), (10,))
    assert approx(J.to_matrix(A), np.diag(diag))
    assert state(A, "diagonal") is diag
    B = JopFooBar(5, g)
    assert approx(J.to_matrix(B), np.diag(state(B)["A"].reshape(-1, order="F")))
    m, d = J.rand(domain(B), g), J.rand(range_(B), g)
    assert approx(jacobian(B, J.rand(domain(B), g)) * m, B * m)
    assert approx(jacobian_(B, J.rand(domain(B), g)) * m, B * m)
    assert approx(adjoint(jacobian(B.T, J.rand(domain(B.T), g))) * d, B.T * d)


def test_nonlinear_operator_and_linearizations(g):  # :170-217
    F = JopBar(10)
    m = J.rand(domain(F), g)
    assert np.array_equal(F * m, m ** 2)
    Jc = jacobian_(F, m)
    assert J.point(Jc) is m
    d = Jc * m
    assert np.array_equal(d, 2 * m * m)
    assert np.array_equal(Jc.T * d, 2 * m * d)
    # upstate (:196-201)
    R = JopRosenbrock()
    mm = g.random(2)
    JR = jacobian_(R, mm)
    assert approx(state(JR)["J"], np.array([[-1.0, 0.0], [-20 * mm[0], 10.0]]))
    # multiple simultaneous linearizations (:203-217)
    F = JopBar(2)
    J1 = jacobian(F, np.array([1.0, 2.0]))
    J2 = jacobian(F, np.array([3.0, 4.0]))
    dm = np.array([1.0, 2.0])
    assert np.array_equal(J1 * dm, 2 * np.array([1.0, 2.0]) * dm)
    assert np.array_equal(J2 * dm, 2 * np.array([3.0, 4.0]) * dm)
    J1 = jacobian_(F, np.array([1.0, 2.0]))
    J2 = jacobian_(F, np.array([3.0, 4.0]))
    assert np.array_equal(J1 * dm, 2 * np.array([3.0, 4.0]) * dm)  # aliasing
    assert np.array_equal(J2 * dm, J1 * dm)
    with pytest.raises(TypeError):
        adjoint(F)


def test_composition_linear(g):  # :296-356
    B = [g.random((10, 10)) for _ in range(4)]
    A1, A2, A3, A4 = [JopBaz(b) for b in B]
    A21 = compose(A2, A1)
    A321 = compose(compose(A3, A2), A1)
    A4321 = A4 @ A3 @ A2 @ A1
    m = J.rand(domain(A1), g)
    assert approx(A21 * m, B[1] @ (B[0] @ m))
    assert approx(A321 * m, B[2] @ (B[1] @ (B[0] @ m)))
    d = A4321 * m
    assert approx(d, B[3] @ (B[2] @ (B[1] @ (B[0] @ m))))
    assert len(state(A4321)["ops"]) == 4  # flattening, :309
    assert approx(A21.T * d, B[0].T @ (B[1].T @ d))
    assert approx(A4321.T * d, B[0].T @ (B[1].T @ (B[2].T @ (B[3].T @ d))))
    assert domain(A4321) == JetSpace(np.float64, 10)
    assert approx(J.to_matrix(A4321) @ m, A4321 * m)
    C = A4 @ A3 @ A21.T  # :324-325
    assert approx(C * m, A4 * (A3 * (A1.T * (A2.T * m))))
    # with a raw matrix (:328-356)
    A3m = B[2]
    A4321m = A4 @ (A3m @ (A2 @ A1))
    assert approx(A4321m * m, d)
    assert approx(A4321m.T * d, A4321.T * d)


def test_composition_nonlinear(g):  # :358-390
    F1, F2, F3, F4 = [JopBar(10) for _ in range(4)]
    F21, F321, F4321 = F2 @ F1, F3 @ F2 @ F1, F4 @ F3 @ F2 @ F1
    m = J.rand(domain(F1), g)
    assert approx(F21 * m, F2 * (F1 * m))
    assert approx(F4321 * m, F4 * (F3 * (F2 * (F1 * m))))
    m = np.ones(10)
    J1 = jacobian_(F1, m)
    J21 = jacobian_(F2, F1 * m) @ J1
    J4321 = (jacobian_(F4, (F3 @ F2 @ F1) * m) @ jacobian_(F3, (F2 @ F1) * m)
             @ jacobian_(F2, F1 * m) @ J1)
    L21 = jacobian_(F21, m)
    L4321 = jacobian_(F4321, m)
    dm = np.ones(10)
    assert approx(J21 * dm, L21 * dm)
    assert approx(J4321 * dm, L4321 * dm)
    dd = J4321 * dm
    assert approx(J4321.T * dd, L4321.T * dd)


def test_composition_mixed_adjoints(g):  # :392-423
    A2 = JopBaz(g.random((10, 10)))
    A4 = JopFoo(g.random(10))
    F1, F3 = JopBar(10), JopBar(10)
    F4321 = A4 @ F3 @ A2.T @ F1
    m = J.rand(domain(F1), g)
    assert approx(F4321 * m, A4 * (F3 * (A2.T * (F1 * m))))
    m = g.random(10)
    J4321 = A4 @ jacobian_(F3, A2.T * (F1 * m)) @ A2.T @ jacobian_(F1, m)
    L4321 = jacobian_(F4321, m)
    dm = g.random(10)
    assert approx(J4321 * dm, L4321 * dm)


def test_composition_block_getblock_and_state(g):  # :425-451
    A1 = JopFoo(g.random(2))
    A2 = J.blockop([JopBar(2), JopBar(2)])
    A = A2 @ A1
    m = J.rand(domain(A), g)
    assert approx(J.getblock(A * m, 1), J.getblock(A, 1, 1) * m)
    assert approx(J.getblock(A * m, 2), J.getblock(A, 2, 1) * m)
    diag = g.random((2, 2))
    Ad, F = JopFoo(diag), JopBar(4)
    G = Ad @ F
    assert state(G, "diagonal") is diag
    with pytest.raises(KeyError):
        state(G, "foo")
    with pytest.raises(KeyError):
        state(Ad @ JopFoo(diag), "diagonal")


def test_sums(g):  # :453-510
    B = [g.random((10, 10)) for _ in range(3)]
    A1, A2, A3 = [JopBaz(b) for b in B]
    A12 = A1 + A2
    A123 = A1 + A2 - A3
    m = J.rand(domain(A1), g)
    assert approx(A12 * m, A1 * m + A2 * m)
    assert approx(A123 * m, A1 * m + A2 * m - A3 * m)
    A123 = A12 + A3
    A12312 = A123 - A12  # sign flipping, :464-465
    assert state(A12312)["sgns"] == (1, 1, 1, -1, -1)
    assert approx(A12312 * m, A1 * m + A2 * m + A3 * m - A1 * m - A2 * m)
    d = J.rand(range_(A1), g)
    assert approx(A123.T * d, A1.T * d + A2.T * d + A3.T * d)
    a1, a2, a3 = g.random(3)  # :471-488
    S = a1 * A1 + a2 * A2 - a3 * A3
    assert approx(S * m, a1 * (A1 * m) + a2 * (A2 * m) - a3 * (A3 * m))
    assert approx(S.T * d, a1 * (A1.T * d) + a2 * (A2.T * d) - a3 * (A3.T * d))
    assert approx((A1 + B[1] - A3) * m, A1 * m + B[1] @ m - A3 * m)  # :490-498
    F2 = JopBar(10)  # :500-510
    F12 = A1 + F2
    assert isinstance(F12, JopNl)
    assert approx(F12 * m, A1 * m + F2 * m)
    J12 = jacobian(F12, m)
    assert approx(J12 * m, A1 * m + jacobian(F2, m) * m)


def test_block_arrays(g):  # :512-551
    R = JetBSpace([JetSpace(np.float64, 2), JetSpace(np.float64, 2, 2), JetSpace(np.float64, 2, 3)])
    assert R.indices == [(1, 2), (3, 6), (7, 12)]
    x = J.ones(R)
    assert J.getblock(x, 2).shape == (2, 2)
    J.setblock_(x, 1, math.pi)
    J.setblock_(x, 2, 2 * math.pi)
    J.setblock_(x, 3, 3 * math.pi * np.ones((2, 3)))
    assert approx(J.getblock_(x, 2, np.empty((2, 2))), 2 * math.pi * np.ones((2, 2)))
    _x = J.to_array(x)
    assert np.isclose(J.norm(x), np.linalg.norm(_x), rtol=1e-14)
    assert J.norm(x, 0) == np.count_nonzero(_x)
    assert J.norm(x, math.inf) == np.max(np.abs(_x))
    x = J.randn(R, g)
    _x = J.to_array(x)
    assert J.extrema(x) == (_x.min(), _x.max())
    Rc = JetBSpace([JetSpace(np.complex128, 2), JetSpace(np.complex128, 2, 2)])
    xc = J.rand(Rc, g)
    assert abs(xc).dtype == np.float64
    assert Rc == J.space(J.rand(Rc, g))
    assert np.isclose(J.dot(xc, xc), np.vdot(J.to_array(xc), J.to_array(xc)))


def test_block_arrays_broadcasting(g):  # :553-600
    R = JetBSpace([JetSpace(np.float64, 2), JetSpace(np.float64, 2, 2), JetSpace(np.float64, 2, 3)])
    u, v, w = J.rand(R, g), J.rand(R, g), J.rand(R, g)
    a, b, c = g.random(3)
    x = a * u + b * v + c * w
    assert isinstance(x, J.BlockArray)
    for xi, ui, vi, wi in zip(x.arrays, u.arrays, v.arrays, w.arrays):
        assert np.array_equal(xi, a * ui + b * vi + c * wi)
    y = J.zeros(R).assign(x)
    assert np.array_equal(J.to_array(y), J.to_array(x))
    J.fill_(x, 3.14)
    assert all(x[i] == 3.14 for i in range(1, len(x) + 1))
    xi = g.integers(0, 100, size=len(R)).astype(np.int32)
    y = J.rand(R, g)
    z = J.bmap(lambda p, q: p * q, xi, y)
    assert isinstance(z, J.BlockArray)
    assert np.array_equal(J.to_array(z), xi * J.to_array(y))


def test_block_array_reshaped_from_array(g):  # :602-620
    x = np.asfortranarray(g.random((5, 10)))
    R = JetBSpace([JetSpace(np.float64, 5) for _ in range(10)])
    _x = J.reshape(x, R)
    flat = x.reshape(-1, order="F")
    for i in range(1, len(_x) + 1):
        assert _x[i] == flat[i - 1]
    _x.arrays[3][2] = -7.0  # shares memory
    assert x[2, 3] == -7.0
    A = [g.random((10, 10)) for _ in range(5)]
    _A = J.blockop([JopBaz(a) for a in A])
    m = J.rand(domain(_A), g)
    _y = _A * m
    for i in range(5):
        assert approx(J.getblock(_y, i + 1), A[i] @ m)


def test_block_operator(g):  # :622-695
    B11, B13, B14, B21, B23, B24, B32, B33 = [g.random((10, 10)) for _ in range(8)]
    A11, A13, A14, A21, A23, A32, A33 = [JopBaz(b) for b in (B11, B13, B14, B21, B23, B32, B33)]
    A24 = JopBaz(B24).T
    F12, F23, F31 = JopBar(10), JopBar(10), JopBar(10)
    Z22 = J.JopZeroBlock(JetSpace(np.float64, 10), JetSpace(np.float64, 10))
    Z34 = J.JopZeroBlock(JetSpace(np.float64, 10), JetSpace(np.float64, 10))
    assert J.iszero(Z22) and not J.iszero(A11) and not J.iszero(F12)
    C24 = A24 @ JopBar(10)
    F = J.blockop([[A11, F12, A13, A14], [A21, Z22, F23, C24], [F31, A32, A33, Z34]])
    assert isinstance(F, JopNl) and J.isblockop(F)
    assert J.nblocks(F) == (3, 4) and J.nblocks(A11) == (1, 1)
    m = J.rand(domain(F), g)
    d = F * m
    mm, dd = J.to_array(m), J.to_array(d)
    assert approx(dd[0:10], B11 @ mm[0:10] + F12 * mm[10:20] + B13 @ mm[20:30] + B14 @ mm[30:40])
    assert approx(dd[10:20], B21 @ mm[0:10] + F23 * mm[20:30] + C24 * mm[30:40])
    assert approx(dd[20:30], F31 * mm[0:10] + B32 @ mm[10:20] + B33 @ mm[20:30])
    Jc = jacobian_(F, m)
    dm = J.rand(domain(Jc), g)
    ddl = Jc * dm
    J12 = jacobian_(F12, mm[10:20].copy())
    J23 = jacobian_(F23, mm[20:30].copy())
    J24 = jacobian_(C24, mm[30:40].copy())
    J31 = jacobian_(F31, mm[0:10].copy())
    L = J.blockop([[A11, J12, A13, A14], [A21, Z22, J23, J24], [J31, A32, A33, Z34]])
    assert isinstance(L, JopLn)
    assert approx(ddl, L * dm)
    assert approx(L.T * ddl, Jc.T * ddl)
    assert approx(mul_(J.rand(domain(L), g), L.T, ddl), Jc.T * ddl)  # dirty output, :684
    K = J.to_matrix(L)
    assert approx(J.to_array(L * dm), K @ J.to_array(dm))
    _J12 = J.getblock(Jc, 1, 2)
    assert isinstance(_J12, JopLn)
    x = J.rand(domain(J12), g)
    assert approx(J12 * x, _J12 * x)


def test_block_forward_accumulates_into_dirty_output_quirk_Q1(g):
    """src/Jets.jl:1015-1030: _d is never zeroed when ncol>1 -- mul! into a dirty d accumulates."""
    A = J.blockop([[JopFoo(g.random(4)), JopFoo(g.random(4))]])
    m = J.rand(domain(A), g)
    d0 = J.rand(range_(A), g)
    keep = J.to_array(d0).copy()
    mul_(d0, A, m)
    assert approx(J.to_array(d0), keep + J.to_array(A * m))


def test_block_shapes(g):  # :704-787
    Bm = g.random((5, 5))
    B = JopBaz(Bm)
    A = J.blockop([[B]])
    m = J.rand(domain(A), g)
    assert approx(J.to_array(A * m), Bm @ m)
    d = J.rand(range_(A), g)
    assert approx(A.T * d, Bm.T @ J.to_array(d))
    Bs = [g.random((5, 5)) for _ in range(3)]
    A = J.blockop([JopBaz(b) for b in Bs])  # tall and skinny: plain-array domain
    assert isinstance(domain(A), JetSpace) and J.nblocks(A) == (3, 1)
    m = J.rand(domain(A), g)
    assert approx(J.to_array(A * m), np.concatenate([b @ m for b in Bs]))
    d = J.rand(range_(A), g)
    dd = J.to_array(d)
    assert approx(A.T * d, sum(b.T @ dd[5 * i:5 * i + 5] for i, b in enumerate(Bs)))
    G = [JopBar(5) for _ in range(3)]
    F = J.blockop(G)
    Jc = jacobian_(F, m)
    assert approx(J.to_array(Jc * m), np.concatenate([2 * m * m] * 3))
    assert approx(Jc.T * d, sum(2 * m * dd[5 * i:5 * i + 5] for i in range(3)))
    A = J.blockop([[JopBaz(b) for b in Bs]])  # short and fat
    m = J.rand(domain(A), g)
    mm = J.to_array(m)
    assert approx(J.to_array(A * m), sum(b @ mm[5 * i:5 * i + 5] for i, b in enumerate(Bs)))
    d = J.rand(range_(A), g)
    assert approx(J.to_array(A.T * d), np.concatenate([b.T @ J.to_array(d) for b in Bs]))
    # getblock of an adjoint (:760-787)
    Bmat = [[JopBaz(g.random((5, 5))) for _ in range(3)] for _ in range(2)]
    A = J.blockop(Bmat)
    C = A.T
    for i in range(2):
        for j in range(3):
            Cji = J.getblock(C, j + 1, i + 1)
            assert isinstance(Cji, JopAdjoint)
            x = J.rand(domain(Cji), g)
            assert approx(Cji * x, Bmat[i][j].T * x)


def test_scalar_times_operator_and_vec(g):  # :789-838
    A = JopBaz(g.random((10, 10)))
    m = J.rand(domain(A), g)
    assert approx((3.14 * A) * m, 3.14 * (A * m))
    w = np.asfortranarray(g.random((10, 11)))
    A2 = J.JopDiagonal(w)
    x = J.rand(domain(A2), g)
    Bv = J.vec(A2)
    assert len(domain(Bv).size()) == 1 and domain(Bv).size() == (110,)
    assert approx(J.vec(A2 * x), Bv * J.vec(x))


def test_dot_product_linearity_linearization(g):  # :901-930
    A = JopFoo(g.random(10))
    lhs, rhs = J.dot_product_test(A, J.rand(domain(A), g), J.rand(range_(A), g))
    assert np.isclose(lhs, rhs, rtol=1e-14)
    mmask, dmask = J.ones(domain(A)), J.ones(range_(A))
    mmask[0] = 0
    dmask[0] = 0
    lhs, rhs = J.dot_product_test(A, J.rand(domain(A), g), J.rand(range_(A), g), mmask, dmask)
    assert np.isclose(lhs, rhs, rtol=1e-14)
    Ac = JopFoo(g.random(10) + 1j * g.random(10))
    lhs, rhs = J.dot_product_test(Ac, J.rand(domain(Ac), g), J.rand(range_(Ac), g))
    assert np.isclose(lhs, rhs, rtol=1e-14)
    lhs, rhs = J.linearity_test(A, rng=g)
    assert approx(lhs, rhs, 1e-14)
    muobs, muexp = J.linearization_test(JopBar(10), g.random(10), rng=g)
    assert np.isclose(muobs.max(), muexp.max(), rtol=1e-6)


# ---- symmetric spaces: runtests.jl:219-282 ---------------------------------------------------
def indexmap(I):  # runtests.jl:219-225 (1-based index tuple in, index tuple out)
    if I[0] < 5:
        return tuple(I)
    return (I[0] - 4, I[1])


def test_symmetric_space(g):  # :227-258
    R = J.JetSSpace(np.complex128, (8, 4), (4, 4), indexmap)
    assert R.size() == (8, 4) and R.eltype == np.complex128
    assert np.array_equal(J.ones(R).full().real, np.ones((8, 4))) and np.array_equal(J.ones(R).full().imag, np.zeros((8, 4)))
    assert np.array_equal(J.zeros(R).full(), np.zeros((8, 4)))
    assert J.rand(R, g).shape == (8, 4) and J.Array(R).shape == (8, 4) and J.Array(R).dtype == np.complex128
    x = J.rand(R, g)
    z = x.similar()
    assert isinstance(z, J.SymmetricArray) and z.shape == (8, 4)
    y = x.A
    assert np.isclose(J.norm(x), math.sqrt(2 * np.linalg.norm(y) ** 2), rtol=1e-14)
    assert np.isclose(J.norm(x, 2), J.norm(x), rtol=1e-15)
    assert np.isclose(J.norm(x, 1), 2 * np.sum(np.abs(y)), rtol=1e-14)
    assert np.isclose(J.norm(x, math.inf), np.max(np.abs(y)), rtol=1e-15)
    x[1, 1] = 0
    x[6, 1] = 0
    assert J.norm(x, 0) == 2 * np.count_nonzero(y)
    assert J.space(J.rand(R, g)) == R
    assert R.similar((0, 0)).n == (0, 0) and R.similar((0, 0)).M == R.M
    for i in range(1, 33):   # linear indices, column-major over the logical size (:252-257)
        x[i] = i + 1j * i
        assert x[i] == i + 1j * i
    # a mirrored position stores the conjugate in the parent (:470-478)
    x[(7, 2)] = 3 + 4j
    assert x.A[2, 1] == 3 - 4j and x[(7, 2)] == 3 + 4j and x[(3, 2)] == 3 - 4j


@pytest.mark.parametrize("T", [np.float32, np.float64])
def test_symmetric_space_real_eltype(g, T):  # the same identities with a real eltype (conj is the identity)
    R = J.JetSSpace(T, (8, 4), (4, 4), indexmap)
    x = J.rand(R, g)
    y = x.A
    assert y.dtype == np.dtype(T) and x.full().shape == (8, 4)
    assert np.array_equal(x.full()[4:, :], y) and np.array_equal(x.full()[:4, :], y)
    tol = 1e-6 if T == np.float32 else 1e-14
    assert np.isclose(J.norm(x), math.sqrt(2 * np.linalg.norm(y.astype(np.float64)) ** 2), rtol=tol)
    assert np.isclose(J.norm(x, 1), 2 * np.sum(np.abs(y.astype(np.float64))), rtol=tol)
    assert np.isclose(J.norm(x, math.inf), np.max(np.abs(y)), rtol=tol)
    x[6, 1] = 0
    assert J.norm(x, 0) == 2 * np.count_nonzero(y)


def test_symmetric_space_broadcast(g):  # :260-282
    R = J.JetSSpace(np.complex128, (8, 4), (4, 4), indexmap)
    u, v, w = J.rand(R, g), J.rand(R, g), J.rand(R, g)
    a, b, c = g.random(3)
    x = a * u + b * v + c * w
    assert isinstance(x, J.SymmetricArray)
    assert np.array_equal(x.A, a * u.A + b * v.A + c * w.A)
    y = J.zeros(R).assign(x)
    assert np.array_equal(y.A, x.A)


# ---- the build's own primitives, pinned the way SURVEY §8c prescribes for the stencil -------
@pytest.mark.parametrize("kind", ["fdiff", "lap"])
@pytest.mark.parametrize("T", [np.float32, np.float64])
def test_stencil_adjoint_and_matrix(kind, T, g):
    n = 17
    S = J.JopStencil(T, n, kind)
    K = J.to_matrix(S)
    ref = np.zeros((n, n), dtype=T)
    for i in range(n):
        if kind == "fdiff" and i < n - 1:
            ref[i, i], ref[i, i + 1] = -1, 1
        if kind == "lap":
            ref[i, i] = -2
            if i > 0:
                ref[i, i - 1] = 1
            if i < n - 1:
                ref[i, i + 1] = 1
    assert np.array_equal(K, ref)
    assert np.array_equal(J.to_matrix(S.T), ref.T)
    lhs, rhs = J.dot_product_test(S, J.rand(domain(S), g), J.rand(range_(S), g))
    assert np.isclose(lhs, rhs, rtol=1e-5 if T == np.float32 else 1e-13)


@pytest.mark.parametrize("fn", list(J.PW_FUNCS))
def test_pointwise_registry_linearization(fn, g):
    F = J.JopPointwise(np.float64, 12, fn, 2.5)
    mo = 0.5 + g.random(12)
    muobs, muexp = J.linearization_test(F, mo, rng=g)
    assert np.isclose(muobs[-1], 4.0, rtol=0.2)
    Jc = jacobian(F, mo)
    lhs, rhs = J.dot_product_test(Jc, g.random(12), g.random(12))
    assert np.isclose(lhs, rhs, rtol=1e-13)


# ---- JetPack-style leaves and state plumbing ----------------------------
@pytest.mark.parametrize("T", [np.float32, np.float64, np.complex128])
def test_restriction_adjoint_and_matrix(T, g):
    n = 23
    idx = g.permutation(n)[:9] + 1          # 1-based, unique, unsorted
    R = J.JopRestriction(T, n, idx)
    m = J.rand(domain(R), g)
    d = J.rand(range_(R), g)
    assert np.array_equal(R * m, m[idx - 1])
    back = R.T * d
    assert np.array_equal(back[idx - 1], d) and np.count_nonzero(back) <= idx.size
    lhs, rhs = J.dot_product_test(R, m, d)
    assert np.isclose(lhs, rhs, rtol=1e-6 if np.dtype(T) == np.float32 else 1e-13)
    M = J.to_matrix(R)
    E = np.zeros((idx.size, n))
    E[np.arange(idx.size), idx - 1] = 1
    assert np.array_equal(M, E.astype(T))


@pytest.mark.parametrize("fn", ["log", "atan"])
def test_new_pointwise_functions_linearization(fn, g):
    F = J.JopPointwise(np.float64, 12, fn)
    mo = g.random(12) + 0.5
    muobs, muexp = J.linearization_test(F, mo, rng=g)
    assert np.isclose(muobs[-1], muexp[-1], rtol=0.1)
    Jm = J.to_matrix(jacobian(F, mo))
    eps = 1e-6
    fd = (F * (mo + eps) - F * (mo - eps)) / (2 * eps)
    assert np.allclose(np.diag(Jm), fd, rtol=1e-6)


# ---- close / perfstat / blockop keyword arguments: runtests.jl:697-702, :840-899 -----------
def JopClose(diag, files):  # :11-18 -- a leaf that owns a resource (a file) released by close
    import tempfile
    fd, path = tempfile.mkstemp()
    os.close(fd)
    files.append(path)

    def df(d, m, *, diagonal, **kw):
        d[...] = diagonal * m
        return d
    spc = JetSpace(np.float64, diag.size)
    return JopLn(df=df, dom=spc, rng=spc,
                 s={"diagonal": diag, "file": path, "_close": lambda j: os.remove(j.s["file"])})


def test_close_releases_leaf_resources_through_every_combinator(g):  # :840-886
    files = []
    A = JopClose(g.random(2), files)
    assert os.path.isfile(state(A)["file"])
    J.close(A)
    assert not os.path.isfile(state(A)["file"])
    # block operator :847-860
    blocks = [[JopClose(g.random(2), files) for _ in range(2)] for _ in range(2)]
    A = J.blockop(blocks)
    parts = [J.getblock(A, i, j) for i in (1, 2) for j in (1, 2)]
    assert all(os.path.isfile(state(B)["file"]) for B in parts)
    J.close(A)
    assert not any(os.path.isfile(state(B)["file"]) for B in parts)
    # composition :862-873 and linear combination :875-886
    for combine in (lambda a2, a1: compose(a2, a1), lambda a2, a1: a2 + a1):
        A1, A2 = JopClose(g.random(2), files), JopClose(g.random(2), files)
        A = combine(A2, A1)
        assert os.path.isfile(state(A1)["file"]) and os.path.isfile(state(A2)["file"])
        J.close(A)
        assert not os.path.isfile(state(A1)["file"]) and not os.path.isfile(state(A2)["file"])
    assert not any(os.path.exists(f) for f in files)


def test_perfstat_is_looked_up_through_compositions_and_sums(g):  # :888-899
    A1 = JopFoo(g.random(2))
    state(A1)["_perfstat"] = lambda j: math.pi      # Jets.perfstat(::Jet{..JopFoo_df!}) = π, :9
    A2 = JopBar(2)
    A = compose(A2, A1)
    assert J.perfstat(A1) == math.pi
    assert J.perfstat(A2) is None
    assert J.perfstat(A) == math.pi
    assert J.perfstat(A2 + A1) == math.pi


def test_blockop_accepts_keyword_arguments(g):  # :697-702
    x = J.JopBlock(np.array([[JopBaz(g.random((2, 2)))], [JopBaz(g.random((2, 2)))]], dtype=object), foo=3)
    assert isinstance(x, J.Jop)
    x = J.JopBlock(np.array([[JopBaz(g.random((2, 2)))], [JopBaz(g.random((2, 2)))]], dtype=object), foo=3, bar=4)
    assert isinstance(x, J.Jop) and state(x)["foo"] == 3 and state(x)["bar"] == 4


# ---- block operators as blocks of a block operator -------------------------------------------
def test_nested_block_operator_is_broken_in_the_reference(g):
    """JetBSpace accepts any JetAbstractSpace as a block (src/Jets.jl:736-751) and reshape / zeros / norm / dot recurse
    through nested BlockArrays, so a JopBlock whose blocks are JopBlocks can be WRITTEN -- but JetBlock_df! never
    zeroes its output (`_d .+= mul!(dtmp, ...)`, :1024, quirk Q1) and the outer operator hands every inner one the
    same `dtmp`, so the inner row sums pile up across block columns: the forward result is wrong and the
    reference's own dot_product_test fails.  (A single block COLUMN of inner operators is fine: `mul!(_d, op, _m)`
    overwrites, :1026.)  That is why jets_op_block rejects children with block spaces instead of "supporting" them."""
    n = 7
    leaves = {}

    def inner(tag, R, C):
        leaves[tag] = [[J.JopDiagonal(g.random(n)) if (r + c) % 2 == 0 else J.JopStencil(np.float64, n, "fdiff") for c in range(C)]
                       for r in range(R)]
        return J.blockop(leaves[tag])
    A = J.blockop([[inner((r, c), 3, 2) for c in range(2)] for r in range(2)])
    assert J.nblocks(A) == (2, 2) and J.size(A) == (42, 28)
    flat = J.blockop([[leaves[(R, C)][r][c] for C in range(2) for c in range(2)] for R in range(2) for r in range(3)])
    m = g.random(28)
    x = J.reshape(m.copy(), J.domain(A))
    assert isinstance(J.getblock(x, 1), J.BlockArray) and J.to_array(x).tobytes() == m.tobytes()   # nested views of one vector
    y, yf = J.to_array(A * x), J.to_array(flat * J.reshape(m.copy(), J.domain(flat)))
    assert not np.allclose(y, yf)                              # the accumulated dtmp
    lhs, rhs = J.dot_product_test(A, x, A * x)
    assert not np.isclose(lhs, rhs, rtol=1e-6)                 # the reference's own self-test rejects it
    d = g.random(42)
    z = J.to_array(A.T * J.reshape(d.copy(), J.range_(A)))      # the adjoint zeroes `_m` (:1044) and is right
    assert np.allclose(z, J.to_array(flat.T * J.reshape(d.copy(), J.range_(flat))), rtol=1e-13)
    col = J.blockop([[inner(("c", r), 2, 1)] for r in range(2)])   # one block column of inner operators: overwrite, correct
    mc = g.random(n)
    flatc = J.blockop([[leaves[("c", R)][r][0]] for R in range(2) for r in range(2)])
    assert np.array_equal(J.to_array(col * mc), J.to_array(flatc * mc))
