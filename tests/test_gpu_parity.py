"""GPU parity tests: every scenario runs through the numpy oracle and through libjets_b200.so
(via the C ABI) on the same seeded inputs.

Tolerances (BASELINE.json north_star): block indexing/layout bit-exact; Float64 1e-12 relative,
Float32 1e-5 relative.  Elementwise / stencil / block / sum / composite paths are additionally
required to be BIT-IDENTICAL to the oracle (the kernels evaluate the same IEEE operations in the
same order with FMA contraction disabled); reductions and dense products are tolerance-checked.
The scenarios restate the reference's hot-path testsets (test/runtests.jl, lines cited).
"""
import math

import numpy as np
import pytest

from backends import OracleBackend, DeviceBackend

pytestmark = pytest.mark.gpu

TOL = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-12}


@pytest.fixture(scope="module")
def O():
    return OracleBackend()


@pytest.fixture(scope="module")
def D():
    return DeviceBackend()


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    nb = np.linalg.norm(b)
    normwise = np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)
    floor = np.max(np.abs(b)) if b.size else 1.0
    elem = np.max(np.abs(a - b) / np.maximum(np.abs(b), floor * 1e-3 + 1e-300)) if b.size else 0.0
    return max(normwise, elem)


def assert_close(a, b, T):
    e = relerr(a, b)
    assert e <= TOL[np.dtype(T)], f"relative error {e:.3e} > {TOL[np.dtype(T)]:.0e}"


def assert_bits(a, b):
    a, b = np.atleast_1d(np.asarray(a)), np.atleast_1d(np.asarray(b))
    assert a.dtype == b.dtype and a.shape == b.shape, (a.dtype, b.dtype, a.shape, b.shape)
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), \
        f"not bit-identical: max abs diff {np.max(np.abs(a.astype(np.float64) - b.astype(np.float64))):.3e}"


def run_both(scn, O, D, *args):
    return scn(O, *args), scn(D, *args)


# ------------------------------------------------------------------ block arrays ------------
def scn_blockarray(K, T, data):
    R = K.JetBSpace([K.JetSpace(T, 2), K.JetSpace(T, 2, 2), K.JetSpace(T, 2, 3), K.JetSpace(T, 1031)])
    x = K.arr(data["x"], R)
    y = K.arr(data["y"], R)
    out = {"ranges": np.array(K.block_ranges(x), dtype=np.int64)}
    out["x"] = K.host(x)
    out["b3"] = np.asarray(K.host(K.getblock(x, 3))).reshape(-1, order="F")
    for p in (2, 1, 0, math.inf, -math.inf, 3.5):
        out[f"norm{p}"] = np.float64(K.norm(x, p))
    out["dot"] = np.float64(K.dot(x, y))
    mn, mx = K.extrema(x)
    out["mn"], out["mx"] = np.float64(mn), np.float64(mx)
    a, b, c = (float(v) for v in data["abc"])
    out["lin"] = K.host(K.lincomb([(a, x), (b, y), (c, x)]))
    out["had"] = K.host(K.hadamard(x, y))
    K.setblock_(x, 2, K.host(K.getblock(y, 2)))
    K.setblock_(x, 1, 3.25)
    out["set"] = K.host(x)
    z = K.zeros(R)
    K.fill_(z, 3.14)
    out["fill"] = K.host(z)
    return out


@pytest.mark.parametrize("T", [np.float32, np.float64])
def test_block_arrays(O, D, T):  # runtests.jl:512-600
    g = np.random.default_rng(1)
    n = 2 + 4 + 6 + 1031
    x = g.standard_normal(n).astype(T)
    x[5] = 0
    data = {"x": x, "y": g.random(n).astype(T), "abc": g.random(3)}
    o, d = run_both(scn_blockarray, O, D, T, data)
    assert np.array_equal(o["ranges"], d["ranges"])  # bit-exact block indexing
    assert np.array_equal(d["ranges"], [[1, 2], [3, 6], [7, 12], [13, 1043]])
    for k in ("x", "b3", "lin", "had", "set", "fill", "mn", "mx", "norm0", "norminf", "norm-inf"):
        assert_bits(o[k], d[k])
    for k in ("norm2", "norm1", "norm3.5", "dot"):
        assert_close(d[k], o[k], T)


def test_reshape_shares_memory(D):  # runtests.jl:602-610
    B = D.B
    x = B.to_device(np.arange(50, dtype=np.float64))
    R = B.JetBSpace([B.JetSpace(np.float64, 5) for _ in range(10)])
    xb = B.reshape(x, R)
    B.setblock_(xb, 4, -7.0)
    h = x.to_host()
    assert np.all(h[15:20] == -7.0) and h[14] == 14 and h[20] == 20
    with pytest.raises(B.JetsError):
        B.reshape(x, B.JetBSpace([B.JetSpace(np.float64, 7)]))


# ------------------------------------------------------------------ leaves ------------------
def scn_linear(K, T, data):
    n = data["w"].size
    A = K.JopDiagonal(data["w"])
    m = K.arr(data["m"], K.domain(A))
    d = A * m
    a = K.adjoint(A) * d
    d2 = K.zeros(K.range_(A))
    K.mul_(d2, A, m)
    return {"d": K.host(d), "a": K.host(a), "d2": K.host(d2), "K": K.to_matrix(A) if n <= 16 else np.zeros(1, T),
            "size": np.array(K.size(A))}


@pytest.mark.parametrize("T", [np.float32, np.float64])
@pytest.mark.parametrize("n", [10, 4099, 1 << 20])
def test_linear_operator(O, D, T, n):  # runtests.jl:126-151
    g = np.random.default_rng(2)
    data = {"w": g.random(n).astype(T), "m": g.random(n).astype(T)}
    o, d = run_both(scn_linear, O, D, T, data)
    for k in o:
        assert_bits(o[k], d[k])
    assert_bits(d["d"], data["w"] * data["m"])


def scn_nonlinear(K, T, fn, data):
    n = data["m"].size
    F = K.JopPointwise(T, n, fn, 2.5)
    m = K.arr(data["m"], K.domain(F))
    dm = K.arr(data["dm"], K.domain(F))
    out = {"F": K.host(F * m)}
    Jc = K.jacobian_(F, m)
    out["J"] = K.host(Jc * dm)
    out["Jt"] = K.host(K.adjoint(Jc) * dm)
    # multiple simultaneous linearizations (runtests.jl:203-217)
    m2 = K.arr(data["m2"], K.domain(F))
    J1 = K.jacobian(F, m)
    J2 = K.jacobian(F, m2)
    out["J1"] = K.host(J1 * dm)
    out["J2"] = K.host(J2 * dm)
    Ja = K.jacobian_(F, m)
    Jb = K.jacobian_(F, m2)
    out["Ja"] = K.host(Ja * dm)   # aliases Jb: linearized about m2
    out["Jb"] = K.host(Jb * dm)
    return out


@pytest.mark.parametrize("T", [np.float32, np.float64])
@pytest.mark.parametrize("fn", ["square", "power", "exp", "sin", "tanh"])
def test_nonlinear_operator(O, D, T, fn):  # runtests.jl:170-217
    g = np.random.default_rng(3)
    n = 3001
    data = {"m": (0.5 + g.random(n)).astype(T), "dm": g.random(n).astype(T), "m2": (0.5 + g.random(n)).astype(T)}
    o, d = run_both(scn_nonlinear, O, D, T, fn, data)
    for k in o:
        if fn == "square":
            assert_bits(o[k], d[k])
        else:
            assert_close(d[k], o[k], T)
    assert np.array_equal(d["Ja"], d["Jb"]) and not np.array_equal(d["J1"], d["J2"])


def test_jacobian_before_point_fails(D):
    B = D.B
    F = B.JopPointwise(np.float64, 8, "square")
    L = B.JopLn(B.core._lin_handle(F), F.dom, F.rng)
    with pytest.raises(B.JetsError) as e:
        L * B.ones(F.dom)
    assert e.value.code == 7
    with pytest.raises(B.JetsError):
        B.adjoint(F)


@pytest.mark.parametrize("T", [np.float32, np.float64])
@pytest.mark.parametrize("kind", ["fdiff", "lap"])
@pytest.mark.parametrize("n", [1, 2, 17, 2048, 2049, 40001])
def test_stencil(O, D, T, kind, n):  # no reference definition: pinned by matrix + adjoint tests
    g = np.random.default_rng(4)
    m = g.standard_normal(n).astype(T)
    d = g.standard_normal(n).astype(T)

    def scn(K):
        S = K.JopStencil(T, n, kind)
        out = {"f": K.host(S * K.arr(m, K.domain(S))), "t": K.host(K.adjoint(S) * K.arr(d, K.range_(S)))}
        if n <= 17:
            out["K"] = K.to_matrix(S)
            out["Kt"] = K.to_matrix(K.adjoint(S))
        return out
    o, dv = scn(O), scn(D)
    for k in o:
        assert_bits(o[k], dv[k])
    if n <= 17:
        assert np.array_equal(dv["K"].T, dv["Kt"])


# ------------------------------------------------------------------ composition -------------
def scn_chain(K, T, data):
    """config 2: diagonal ∘ finite-difference ∘ jacobian(pointwise JopNl)."""
    n = data["w"].size
    Dg = K.JopDiagonal(data["w"])
    S = K.JopStencil(T, n, "fdiff")
    F = K.JopPointwise(T, n, "square")
    G = Dg @ S @ F                       # nonlinear composite
    mo = K.arr(data["mo"], K.domain(G))
    dm = K.arr(data["dm"], K.domain(G))
    dd = K.arr(data["dd"], K.range_(G))
    out = {"G": K.host(G * mo)}
    Jc = K.jacobian(G, mo)
    out["J"] = K.host(Jc * dm)
    out["Jt"] = K.host(K.adjoint(Jc) * dd)
    A = Dg @ S @ K.jacobian(F, mo)       # the same thing assembled by hand (runtests.jl:371-389)
    out["A"] = K.host(A * dm)
    out["At"] = K.host(K.adjoint(A) * dd)
    lhs, rhs = K.dot_product_test(A, dm, dd)
    out["dpt"] = np.array([lhs, rhs], dtype=np.float64)
    return out


def dot_ref(x, y):
    """f64 reference dot and its condition scale sum|x_i y_i| (the classic dot error bound)."""
    x, y = np.asarray(x, np.float64).ravel(), np.asarray(y, np.float64).ravel()
    return float(x @ y), float(np.abs(x) @ np.abs(y))


@pytest.mark.parametrize("T", [np.float32, np.float64])
@pytest.mark.parametrize("n", [10, 2048 * 3 + 5, 1_000_000])
def test_composite_chain_config2(O, D, T, n):
    g = np.random.default_rng(5)
    data = {k: g.random(n).astype(T) for k in ("w", "mo", "dm", "dd")}
    o, d = run_both(scn_chain, O, D, T, data)
    for k in ("G", "J", "Jt", "A", "At"):
        assert_bits(o[k], d[k])
    assert_bits(d["J"], d["A"])
    # dot_product_test: the vectors are bit-identical to the oracle's, so the two dots are checked
    # against an f64 evaluation, relative to the dot's condition scale
    for got, (ref, scale) in zip(d["dpt"], (dot_ref(data["dm"], d["At"]), dot_ref(d["A"], data["dd"]))):
        assert abs(got - ref) <= TOL[np.dtype(T)] * scale
    assert abs(d["dpt"][0] - d["dpt"][1]) <= TOL[np.dtype(T)] * dot_ref(d["A"], data["dd"])[1]


def test_chain_is_one_fused_launch(D):
    B = D.B
    n = 1 << 16
    g = np.random.default_rng(6)
    A = B.JopDiagonal(g.random(n).astype(np.float32)) @ B.JopStencil(np.float32, n) @ \
        B.jacobian(B.JopPointwise(np.float32, n), B.to_device(g.random(n).astype(np.float32)))
    x = B.rand(B.domain(A))
    y = B.zeros(B.range_(A))
    B.mul_(y, A, x)
    c0 = B.launch_count()
    B.mul_(y, A, x)
    B.mul_(x, A.T, y)
    assert B.launch_count() - c0 == 2
    assert B.plan_info(A)["engines"] == ["tma"] and B.plan_info(A)["launches"] == 1


def scn_comp_dense(K, T, data):
    B = data["B"]
    A1, A2, A3, A4 = [K.JopDense(b) for b in B]
    A4321 = A4 @ A3 @ A2 @ A1
    m = K.arr(data["m"], K.domain(A1))
    d = A4321 * m
    out = {"d": K.host(d), "a": K.host(K.adjoint(A4321) * d)}
    C = A4 @ A3 @ K.adjoint(A2 @ A1)     # runtests.jl:324-325
    out["c"] = K.host(C * m)
    F1, F3 = K.JopPointwise(T, 10), K.JopPointwise(T, 10)
    A5 = K.JopDiagonal(data["w"])
    G = A5 @ F3 @ K.adjoint(A2) @ F1     # runtests.jl:392-423
    out["g"] = K.host(G * m)
    L = K.jacobian_(G, m)
    dm = K.arr(data["dm"], K.domain(G))
    out["l"] = K.host(L * dm)
    out["lt"] = K.host(K.adjoint(L) * dm)
    return out


@pytest.mark.parametrize("T", [np.float32, np.float64])
def test_composition_with_dense(O, D, T):  # runtests.jl:296-423
    g = np.random.default_rng(7)
    data = {"B": [g.random((10, 10)).astype(T) for _ in range(4)], "m": g.random(10).astype(T),
            "dm": g.random(10).astype(T), "w": g.random(10).astype(T)}
    o, d = run_both(scn_comp_dense, O, D, T, data)
    for k in o:
        assert_close(d[k], o[k], T)


# ------------------------------------------------------------------ sums --------------------
def scn_sums(K, T, data):
    n = data["m"].size
    A1, A2, A3 = [K.JopDiagonal(w) for w in data["w"]]
    S = K.JopStencil(T, n, "lap")
    m = K.arr(data["m"], K.domain(A1))
    d = K.arr(data["d"], K.range_(A1))
    A12 = A1 + A2
    A123 = A1 + A2 - A3
    A12312 = (A12 + A3) - A12            # sign flipping, runtests.jl:464-465
    out = {"a": K.host(A12 * m), "b": K.host(A123 * m), "c": K.host(A12312 * m),
           "ct": K.host(K.adjoint(A12312) * d)}
    a1, a2 = (float(v) for v in data["a"])
    Sc = a1 * A1 + a2 * A2 - 0.5 * S      # runtests.jl:471-488 and config 4's B - c*S
    out["s"] = K.host(Sc * m)
    out["st"] = K.host(K.adjoint(Sc) * d)
    F = K.JopPointwise(T, n, "square")
    F12 = A1 + F                          # runtests.jl:500-510
    out["f"] = K.host(F12 * m)
    J12 = K.jacobian(F12, m)
    out["j"] = K.host(J12 * d)
    return out


@pytest.mark.parametrize("T", [np.float32, np.float64])
@pytest.mark.parametrize("n", [10, 70001])
def test_sums(O, D, T, n):
    g = np.random.default_rng(8)
    data = {"w": [g.random(n).astype(T) for _ in range(3)], "m": g.random(n).astype(T),
            "d": g.random(n).astype(T), "a": g.random(2)}
    o, d = run_both(scn_sums, O, D, T, data)
    for k in ("a", "b", "c", "ct", "f", "j"):
        assert_bits(o[k], d[k])
    for k in ("s", "st"):  # scalar cast order (a*A = scale ∘ A): compare at tolerance
        assert_close(d[k], o[k], T)


# ------------------------------------------------------------------ block operators ---------
def scn_block_diag(K, T, data):
    """config 1: R x C diagonal JopLn blocks."""
    W = data["W"]
    nr, nc = len(W), len(W[0])
    A = K.blockop([[K.JopDiagonal(W[r][c]) for c in range(nc)] for r in range(nr)])
    m = K.arr(data["m"], K.domain(A))
    d = K.arr(data["d"], K.range_(A))
    out = {"nb": np.array(K.nblocks(A)), "f": K.host(A * m), "t": K.host(K.adjoint(A) * d)}
    lhs, rhs = K.dot_product_test(A, m, d)
    out["dpt"] = np.array([lhs, rhs], dtype=np.float64)
    dirty = K.arr(data["dirty"], K.domain(A))
    out["tdirty"] = K.host(K.mul_(dirty, K.adjoint(A), d))     # runtests.jl:684
    l1, l2 = K.linearity_test(A, K.arr(data["m"], K.domain(A)), K.arr(data["m2"], K.domain(A)))
    out["lin1"], out["lin2"] = K.host(l1), K.host(l2)
    return out


@pytest.mark.parametrize("T", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(4, 4, 1000), (3, 2, 4099), (1, 3, 513), (3, 1, 6000), (2, 70, 256)])
def test_block_diagonal_config1_small(O, D, T, shape):
    nr, nc, n = shape
    g = np.random.default_rng(9)
    data = {"W": [[g.random(n).astype(T) for _ in range(nc)] for _ in range(nr)],
            "m": g.random(nc * n).astype(T), "m2": g.random(nc * n).astype(T),
            "d": g.random(nr * n).astype(T), "dirty": g.random(nc * n).astype(T)}
    o, d = run_both(scn_block_diag, O, D, T, data)
    assert np.array_equal(o["nb"], d["nb"])
    for k in ("f", "t", "tdirty", "lin1", "lin2"):
        assert_bits(o[k], d[k])
    assert_close(d["dpt"], o["dpt"], T)
    assert_close(d["lin1"], d["lin2"], T)


def test_block_diagonal_config1_full_size(O, D):
    """BASELINE config 1 at full size: 4x4 of 1e6-element diagonal blocks, Float64."""
    T, n = np.float64, 1_000_000
    g = np.random.default_rng(10)
    data = {"W": [[g.random(n) for _ in range(4)] for _ in range(4)], "m": g.random(4 * n), "m2": g.random(4 * n),
            "d": g.random(4 * n), "dirty": g.random(4 * n)}
    o, d = run_both(scn_block_diag, O, D, T, data)
    for k in ("f", "t", "tdirty"):
        assert_bits(o[k], d[k])
    assert_close(d["dpt"], o["dpt"], T)
    assert abs(d["dpt"][0] - d["dpt"][1]) <= 1e-12 * abs(d["dpt"][0] + d["dpt"][1])


def test_forward_block_accumulate_quirk_Q1(O, D):
    """src/Jets.jl:1001,1024: the reference's forward block mul! adds into a dirty d when ncol>1.
    accumulate=True reproduces it; the default overwrites (documented deviation)."""
    T = np.float64
    g = np.random.default_rng(11)
    w = [g.random(300) for _ in range(2)]
    m, d0 = g.random(600), g.random(300)
    Ao = O.blockop([[O.JopDiagonal(w[0]), O.JopDiagonal(w[1])]])
    ref = O.host(O.mul_(O.arr(d0, O.range_(Ao)), Ao, O.arr(m, O.domain(Ao))))
    Ad = D.blockop([[D.JopDiagonal(w[0]), D.JopDiagonal(w[1])]])
    got = D.host(D.B.mul_(D.arr(d0, D.range_(Ad)), Ad, D.arr(m, D.domain(Ad)), accumulate=True))
    assert_bits(ref, got)
    clean = D.host(D.B.mul_(D.arr(d0, D.range_(Ad)), Ad, D.arr(m, D.domain(Ad))))
    assert_bits(clean, O.host(Ao * O.arr(m, O.domain(Ao))))


def scn_block_mixed(K, T, data):
    """runtests.jl:622-695: 3x4 of dense / pointwise / zero / composite / adjoint blocks."""
    Bm = data["B"]
    sp = K.JetSpace(T, 10)
    A11, A13, A14, A21, A23, A32, A33 = [K.JopDense(Bm[i]) for i in range(7)]
    A24 = K.adjoint(K.JopDense(Bm[7]))
    F12, F23, F31 = K.JopPointwise(T, 10), K.JopPointwise(T, 10), K.JopPointwise(T, 10)
    Z22, Z34 = K.JopZeroBlock(sp, sp), K.JopZeroBlock(sp, sp)
    C24 = A24 @ K.JopPointwise(T, 10)
    F = K.blockop([[A11, F12, A13, A14], [A21, Z22, F23, C24], [F31, A32, A33, Z34]])
    out = {"iszero": np.array([K.iszero(Z22), K.iszero(A11), K.iszero(F12)]),
           "nb": np.array(K.nblocks(F))}
    m = K.arr(data["m"], K.domain(F))
    out["F"] = K.host(F * m)
    Jc = K.jacobian_(F, m)
    dm = K.arr(data["dm"], K.domain(Jc))
    dd = Jc * dm
    out["J"] = K.host(dd)
    out["Jt"] = K.host(K.adjoint(Jc) * dd)
    out["Jt_dirty"] = K.host(K.mul_(K.arr(data["dirty"], K.domain(Jc)), K.adjoint(Jc), dd))
    out["K"] = K.to_matrix(Jc)
    J12 = K.getblock(Jc, 1, 2)
    x = K.arr(data["x"], K.domain(J12))
    out["J12"] = K.host(J12 * x)
    return out


@pytest.mark.parametrize("T", [np.float32, np.float64])
def test_block_operator_mixed(O, D, T):
    g = np.random.default_rng(12)
    data = {"B": [g.random((10, 10)).astype(T) for _ in range(8)], "m": g.random(40).astype(T),
            "dm": g.random(40).astype(T), "dirty": g.random(40).astype(T), "x": g.random(10).astype(T)}
    o, d = run_both(scn_block_mixed, O, D, T, data)
    assert np.array_equal(o["iszero"], d["iszero"]) and np.array_equal(o["nb"], d["nb"])
    for k in ("F", "J", "Jt", "Jt_dirty", "K", "J12"):
        assert_close(d[k], o[k], T)
    # L*dm ≈ K*dm (runtests.jl:687-689)
    assert_close(d["K"] @ data["dm"].astype(np.float64), d["J"], T)


def scn_block_shapes(K, T, data):
    """runtests.jl:704-787: singleton, tall-and-skinny (plain-array domain), short-and-fat."""
    Bs = data["B"]
    out = {}
    A = K.blockop([[K.JopDense(Bs[0])]])
    m = K.arr(data["m5"], K.domain(A))
    d = K.arr(data["d5"], K.range_(A))
    out["single_f"], out["single_t"] = K.host(A * m), K.host(K.adjoint(A) * d)
    A = K.blockop([K.JopDense(b) for b in Bs])           # tall
    out["tall_nb"] = np.array(K.nblocks(A))
    d15 = K.arr(data["d15"], K.range_(A))
    out["tall_f"] = K.host(A * K.arr(data["m5"], K.domain(A)))
    out["tall_t"] = np.asarray(K.host(K.adjoint(A) * d15)).reshape(-1)
    G = K.blockop([K.JopPointwise(T, 5) for _ in range(3)])
    m5 = K.arr(data["m5"], K.domain(G))
    out["tallnl_f"] = K.host(G * m5)
    Jc = K.jacobian_(G, m5)
    out["tallnl_j"] = K.host(Jc * m5)
    out["tallnl_jt"] = np.asarray(K.host(K.adjoint(Jc) * d15)).reshape(-1)
    A = K.blockop([[K.JopDense(b) for b in Bs]])         # fat
    m15 = K.arr(data["m15"], K.domain(A))
    out["fat_f"] = K.host(A * m15)
    out["fat_t"] = K.host(K.adjoint(A) * K.arr(data["d5"], K.range_(A)))
    # getblock of an adjoint (runtests.jl:760-771)
    A = K.blockop([[K.JopDense(Bs[0]), K.JopDense(Bs[1]), K.JopDense(Bs[2])],
                   [K.JopDense(Bs[2]), K.JopDense(Bs[0]), K.JopDense(Bs[1])]])
    C = K.adjoint(A)
    C32 = K.getblock(C, 3, 2)
    out["C32"] = K.host(C32 * K.arr(data["m5"], K.domain(C32)))
    return out


@pytest.mark.parametrize("T", [np.float32, np.float64])
def test_block_shapes(O, D, T):
    g = np.random.default_rng(13)
    data = {"B": [g.random((5, 5)).astype(T) for _ in range(3)], "m5": g.random(5).astype(T),
            "d5": g.random(5).astype(T), "m15": g.random(15).astype(T), "d15": g.random(15).astype(T)}
    o, d = run_both(scn_block_shapes, O, D, T, data)
    assert np.array_equal(d["tall_nb"], [3, 1])
    for k in o:
        assert_close(d[k], o[k], T)
    assert_close(d["C32"], data["B"][1].astype(np.float64).T @ data["m5"], T)


def test_composite_of_block_getblock(O, D):  # runtests.jl:425-436
    T = np.float64
    g = np.random.default_rng(14)
    w, m = g.random(64), g.random(64)

    def scn(K):
        A = K.blockop([K.JopPointwise(T, 64), K.JopPointwise(T, 64, "exp")]) @ K.JopDiagonal(w)
        x = K.arr(m, K.domain(A))
        y = A * x
        return {"y": K.host(y), "a11": K.host(K.getblock(A, 1, 1) * x), "a21": K.host(K.getblock(A, 2, 1) * x)}
    o, d = scn(O), scn(D)
    for k in o:
        assert_close(d[k], o[k], T)
    assert_bits(d["y"][:64], d["a11"])


# ------------------------------------------------------------------ dense blocks (config 3) --
@pytest.mark.parametrize("T", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(2, 3, 256, 128), (3, 3, 100, 77), (1, 2, 1031, 515), (2, 2, 2048, 2048)])
def test_dense_block_gemv(O, D, T, shape):
    nr, nc, rows, cols = shape
    g = np.random.default_rng(15)
    Bm = [[g.random((rows, cols)).astype(T) for _ in range(nc)] for _ in range(nr)]
    m = g.random(nc * cols).astype(T)
    d = g.random(nr * rows).astype(T)

    def scn(K):
        A = K.blockop([[K.JopDense(Bm[r][c]) for c in range(nc)] for r in range(nr)])
        lhs, rhs = K.dot_product_test(A, K.arr(m, K.domain(A)), K.arr(d, K.range_(A)))
        return {"f": K.host(A * K.arr(m, K.domain(A))), "t": K.host(K.adjoint(A) * K.arr(d, K.range_(A))),
                "dpt": np.array([lhs, rhs], dtype=np.float64)}
    o, dv = scn(O), scn(D)
    M = np.block([[b.astype(np.float64) for b in row] for row in Bm])
    assert_close(dv["f"], M @ m.astype(np.float64), T)
    assert_close(dv["t"], M.T @ d.astype(np.float64), T)
    assert relerr(dv["f"], M @ m) <= relerr(o["f"], M @ m) * 4 + TOL[np.dtype(T)] * 0.1
    assert abs(dv["dpt"][0] - dv["dpt"][1]) <= TOL[np.dtype(T)] * abs(dv["dpt"][0] + dv["dpt"][1])
    Ad = D.blockop([[D.JopDense(Bm[r][c]) for c in range(nc)] for r in range(nr)])
    Ad * D.arr(m, D.domain(Ad))
    info = D.B.plan_info(Ad)
    assert info["engines"] == ["gemv"] and info["launches"] == 1


def test_dense_multi_rhs(O, D):
    T = np.float32
    g = np.random.default_rng(16)
    A = g.random((96, 64)).astype(T)
    M = g.random((64, 8)).astype(T)
    op = D.B.JopDense(A, nrhs=8)
    y = op * D.B.to_device(M, D.B.domain(op))
    assert y.shape == (96, 8)
    assert_close(y.to_host(), A.astype(np.float64) @ M.astype(np.float64), T)
    x = op.T * y
    assert_close(x.to_host(), A.astype(np.float64).T @ y.to_host().astype(np.float64), T)


@pytest.mark.parametrize("shape", [(1, 1, 128, 64, 16), (1, 1, 96, 64, 8), (1, 1, 300, 200, 5), (2, 3, 256, 128, 64),
                                   (3, 2, 100, 76, 33), (1, 2, 1031, 516, 2), (2, 2, 2048, 2048, 64), (1, 1, 64, 4096, 70)])
def test_dense_multi_rhs_tcgen05(O, D, shape):
    """Dense blocks applied to several right-hand sides run on the tcgen05/TMEM path (split-TF32,
    three MMAs per product) and must still meet the Float32 tolerance in both orientations, including
    ragged tiles (rows/cols not multiples of the 128x32 tile), padded right-hand-side counts and more
    than 64 right-hand sides (two launches)."""
    nr, nc, rows, cols, nrhs = shape
    T = np.float32
    B = D.B
    g = np.random.default_rng(160 + rows + nrhs)
    Bm = [[(g.random((rows, cols)) - 0.3).astype(T) for _ in range(nc)] for _ in range(nr)]
    X = [(g.random((cols, nrhs)) - 0.5).astype(T) for _ in range(nc)]
    Y = [(g.random((rows, nrhs)) - 0.5).astype(T) for _ in range(nr)]
    A = B.blockop([[B.JopDense(Bm[r][c], nrhs=nrhs) for c in range(nc)] for r in range(nr)])
    x = B.to_device(np.concatenate([v.reshape(-1, order="F") for v in X]), B.domain(A))
    y = B.to_device(np.concatenate([v.reshape(-1, order="F") for v in Y]), B.range_(A))
    f = (A * x).to_host()
    t = (B.adjoint(A) * y).to_host()
    # TMA needs a 16-byte row pitch: a column-major block with rows % 4 != 0 falls back to one GEMV per column
    assert B.plan_info(A)["engines"] == (["tcgen05"] if rows % 4 == 0 else ["gemv"]), B.plan_info(A)
    f = [f[r * rows * nrhs:(r + 1) * rows * nrhs].reshape((rows, nrhs), order="F") for r in range(nr)]
    t = [t[c * cols * nrhs:(c + 1) * cols * nrhs].reshape((cols, nrhs), order="F") for c in range(nc)]
    # The data is signed, so single outputs can cancel to ~0: the elementwise criterion is the componentwise
    # bound |err_i| <= tol * sum_k |a_ik||x_k| (what "1e-5 relative" means for an inner product), plus the
    # norm-wise relative error Julia's isapprox uses.
    def check(got, terms):
        ref = sum(a @ b for a, b in terms)
        scale = sum(np.abs(a) @ np.abs(b) for a, b in terms)
        assert np.linalg.norm(got - ref) <= TOL[np.dtype(T)] * np.linalg.norm(ref)
        assert np.all(np.abs(got - ref) <= TOL[np.dtype(T)] * scale), float(np.max(np.abs(got - ref) / scale))
    for r in range(nr):
        check(f[r], [(Bm[r][c].astype(np.float64), X[c].astype(np.float64)) for c in range(nc)])
    for c in range(nc):
        check(t[c], [(Bm[r][c].astype(np.float64).T, Y[r].astype(np.float64)) for r in range(nr)])
    lhs, rhs = B.dot_product_test(A, x, y)
    big = float(np.abs(np.concatenate([v.ravel() for v in Y])).sum()) * max(float(np.abs(b).max()) for row in Bm for b in row)
    assert abs(lhs - rhs) <= TOL[np.dtype(T)] * max(abs(lhs), abs(rhs), 1e-3 * big)


def test_dense_multi_rhs_accumulates_with_other_terms(O, D):
    """A block row mixing a tensor-core dense block with a diagonal block: the fused launch SETs the row,
    the GEMM accumulates on top (plan = fused + tcgen05)."""
    T = np.float32
    B = D.B
    g = np.random.default_rng(171)
    n, nrhs = 256, 16
    Am = g.random((n, n)).astype(T)
    w = g.random(n * nrhs).astype(T)
    x = g.random(2 * n * nrhs).astype(T)
    A = B.blockop([[B.JopDense(Am, nrhs=nrhs), B.JopDiagonal(B.to_device(w, B.JetSpace(T, n, nrhs)))]])
    y = (A * B.to_device(x, B.domain(A))).to_host()
    X0 = x[:n * nrhs].reshape((n, nrhs), order="F").astype(np.float64)
    ref = (Am.astype(np.float64) @ X0).reshape(-1, order="F") + w.astype(np.float64) * x[n * nrhs:]
    assert_close(y, ref, T)
    assert "tcgen05" in B.plan_info(A)["engines"]


# ------------------------------------------------------------------ engines, edge cases -----
@pytest.mark.parametrize("T", [np.float32, np.float64])
def test_tma_and_ldg_engines_agree_bitwise(D, T):
    B = D.B
    g = np.random.default_rng(17)
    n = 50_000
    W = [[g.random(n).astype(T) for _ in range(3)] for _ in range(2)]
    m = g.random(3 * n).astype(T)
    res = {}
    for eng in ("tma", "ldg"):
        B.set_fused_engine(eng)
        try:
            A = B.blockop([[B.JopDiagonal(W[r][c]) @ B.JopStencil(T, n, "lap") for c in range(3)] for r in range(2)])
            x = B.to_device(m, B.domain(A))
            res[eng] = ((A * x).to_host(), B.plan_info(A)["engines"])
        finally:
            B.set_fused_engine("auto")
    assert res["tma"][1] == ["tma"] and res["ldg"][1] == ["ldg"]
    assert_bits(res["tma"][0], res["ldg"][0])


def test_unaligned_blocks_and_wrapped_memory(O, D):
    """Odd block lengths put block starts off 16-byte boundaries; caller-owned (torch) memory has no
    guard padding.  Both must fall back to the guarded-load engine and still be bit-exact."""
    import torch
    B = D.B
    T = np.float32
    g = np.random.default_rng(18)
    lens = [3, 1, 1025, 7]
    W = [[g.random(l).astype(T)] for l in lens]

    def scn(K):
        A = K.blockop([[K.JopDiagonal(W[i][0]) if i == j else K.JopZeroBlock(K.JetSpace(T, lens[j]), K.JetSpace(T, lens[i]))
                        for j in range(4)] for i in range(4)])
        m = K.arr(np.concatenate([w[0] for w in W]) + 1, K.domain(A))
        return K.host(A * m), K.host(K.adjoint(A) * m)
    o, d = scn(O), scn(D)
    assert_bits(o[0], d[0])
    assert_bits(o[1], d[1])
    t = torch.arange(4096, dtype=torch.float32, device="cuda")
    x = B.wrap_torch(t)
    A = B.JopDiagonal(np.full(4096, 2.0, dtype=T))
    y = B.wrap_torch(torch.zeros(4096, dtype=torch.float32, device="cuda"))
    B.mul_(y, A, x)
    B.sync()
    assert torch.equal(y._owner, t * 2)
    assert "ldg" in B.plan_info(A)["engines"]
    # operator STATE in caller-owned (unguarded) memory under a stencil chain, applied to library-owned vectors:
    # the TMA engines would fetch 16 bytes before and up to 31 bytes past the caller's allocation, so
    # the planner must route the whole launch to the guarded-load engine -- and the result must not change
    n = 4096
    wt = torch.rand(n, dtype=torch.float32, device="cuda")
    S = B.JopStencil(T, n, "lap")
    Aw = S @ B.JopDiagonal(B.wrap_torch(wt))              # S(w .* x): the stencil reads w[i-1], w[i+1]
    Ag = S @ B.JopDiagonal(wt.cpu().numpy())              # the same operator with library-owned state
    xin = B.rand(B.JetSpace(T, n), seed=5)
    assert_bits((Aw * xin).to_host(), (Ag * xin).to_host())
    assert_bits((Aw.T * xin).to_host(), (Ag.T * xin).to_host())
    assert B.plan_info(Aw)["engines"] == ["ldg"] and B.plan_info(Ag)["engines"] == ["tma"]


def test_shape_and_dtype_errors(D):
    B = D.B
    A = B.JopDiagonal(np.ones(8))
    with pytest.raises(B.JetsError) as e:
        B.mul_(B.zeros(B.JetSpace(np.float64, 9)), A, B.ones(A.dom))
    assert e.value.code == 2
    with pytest.raises(B.JetsError) as e:
        B.mul_(B.zeros(B.JetSpace(np.float32, 8)), A, B.ones(A.dom))
    assert e.value.code == 3
    with pytest.raises(B.JetsError):
        B.compose(A, B.JopDiagonal(np.ones(9)))
    with pytest.raises(B.JetsError) as e:
        B.zeros(B.JetSpace(np.int32, 4))          # eltypes outside Float32/64, ComplexF32/64
    assert e.value.code == 3
    with pytest.raises(B.JetsError) as e:         # a complex vector through a real operator
        B.mul_(B.zeros(B.JetSpace(np.complex128, 8)), A, B.ones(B.JetSpace(np.complex128, 8)))
    assert e.value.code == 3


def test_device_rand_is_partition_invariant(D):
    B = D.B
    R1 = B.JetSpace(np.float64, 10_000)
    R2 = B.JetBSpace([B.JetSpace(np.float64, 2_500) for _ in range(4)])
    a = B.rand(R1, seed=7).to_host()
    b = B.rand(R2, seed=7).to_host()
    assert np.array_equal(a, b) and 0.0 <= a.min() and a.max() < 1.0 and abs(a.mean() - 0.5) < 0.02
    z = B.randn(R1, seed=8).to_host()
    assert abs(z.mean()) < 0.05 and abs(z.std() - 1.0) < 0.05


# ------------------------------------------------------------------ LSQR loop (config 4) ----
def _lsqr_oracle(J, A, b, iters):
    At = J.adjoint(A)
    x = J.zeros(J.domain(A))
    T = b.dtype.type
    u = b.copy()
    beta = float(J.norm(u))
    u = T(1.0 / beta) * u
    v = At * u
    alpha = float(J.norm(v))
    v = T(1.0 / alpha) * v
    w = v.copy()
    phibar, rhobar = beta, alpha
    hist = []
    for _ in range(iters):
        u = T(1.0) * (A * v) + T(-alpha) * u
        beta = float(J.norm(u))
        u = T(1.0 / beta) * u
        v = T(1.0) * (At * u) + T(-beta) * v
        alpha = float(J.norm(v))
        v = T(1.0 / alpha) * v
        rho = (rhobar * rhobar + beta * beta) ** 0.5
        c, s = rhobar / rho, beta / rho
        theta = s * alpha
        rhobar = -c * alpha
        phi = c * phibar
        phibar = s * phibar
        x = T(1.0) * x + T(phi / rho) * w
        w = T(1.0) * v + T(-theta / rho) * w
        hist.append((alpha, beta))
    return x, hist


def test_lsqr_loop_config4(O, D):
    """A = B - 0.5*S (block-diagonal diagonals minus block-diagonal stencils), Float64."""
    T = np.float64
    nb, n = 4, 4096
    g = np.random.default_rng(19)
    W = [1.0 + g.random(n) for _ in range(nb)]
    rhs = g.random(nb * n)

    def build(K):
        sp = K.JetSpace(T, n)
        Bd = K.blockop([[K.JopDiagonal(W[i]) if i == j else K.JopZeroBlock(sp, sp) for j in range(nb)] for i in range(nb)])
        Sd = K.blockop([[K.JopStencil(T, n, "lap") if i == j else K.JopZeroBlock(sp, sp) for j in range(nb)] for i in range(nb)])
        return Bd - 0.5 * Sd
    Ao, Ad = build(O), build(D)
    iters = 25
    xo, ho = _lsqr_oracle(O.J, Ao, O.arr(rhs, O.range_(Ao)), iters)
    xd, hd = D.B.solvers.lsqr(Ad, D.arr(rhs, D.range_(Ad)), iters)
    for (a1, b1), (a2, b2) in zip(ho, hd):
        assert abs(a1 - a2) <= 1e-11 * abs(a1) and abs(b1 - b2) <= 1e-11 * abs(b1)
    assert relerr(D.host(xd), O.host(xo)) <= 1e-9
    xg, (ag, bg) = D.B.solvers.lsqr_graph(Ad, D.arr(rhs, D.range_(Ad)), iters)
    assert relerr(D.host(xg), O.host(xo)) <= 1e-9
    assert abs(ag - ho[-1][0]) <= 1e-10 * abs(ag)
    info = D.B.plan_info(Ad)
    assert info["engines"] == ["tma"] and info["launches"] == 1  # sum + blocks fused into one launch


def test_lsqr_fused_updates_config4(O, D):
    """LsqrGraphFused (u, v kept unnormalised, updates folded into the applies through
    jets_apply_axpby, scalar recurrences in two scalar programs) reproduces the oracle's LSQR."""
    T = np.float64
    nb, n = 4, 4096
    g = np.random.default_rng(19)
    W = [1.0 + g.random(n) for _ in range(nb)]
    rhs = g.random(nb * n)

    def build(K):
        sp = K.JetSpace(T, n)
        Bd = K.blockop([[K.JopDiagonal(W[i]) if i == j else K.JopZeroBlock(sp, sp) for j in range(nb)] for i in range(nb)])
        Sd = K.blockop([[K.JopStencil(T, n, "lap") if i == j else K.JopZeroBlock(sp, sp) for j in range(nb)] for i in range(nb)])
        return Bd - 0.5 * Sd
    Ao, Ad = build(O), build(D)
    iters = 25
    xo, ho = _lsqr_oracle(O.J, Ao, O.arr(rhs, O.range_(Ao)), iters)
    G = D.B.solvers.LsqrGraphFused(Ad, D.arr(rhs, D.range_(Ad)))
    G.run(iters - 1)
    xg, (ag, bg) = G.result()
    assert relerr(D.host(xg), O.host(xo)) <= 1e-9
    assert abs(ag - ho[-1][0]) <= 1e-10 * abs(ag) and abs(bg - ho[-1][1]) <= 1e-10 * abs(bg)


@pytest.mark.parametrize("kind", ["elementwise", "dense"])
def test_apply_axpby_matches_separate_ops(D, kind):
    """jets_apply_axpby: out = cA*(A x) + cO*out, fused into the store epilogue of the block-apply
    kernel (elementwise operator) or staged (dense operator) -- bit-identical to apply + axpby."""
    import ctypes as C
    B = D.B
    L = B.solvers.L
    T = np.float64
    g = np.random.default_rng(23)
    n, nb = 6000, 3
    if kind == "elementwise":
        W = [[g.random(n) for _ in range(nb)] for _ in range(nb)]
        A = B.blockop([[B.JopDiagonal(W[r][c]) if r != c else 0.5 * B.JopStencil(T, n, "lap") for c in range(nb)] for r in range(nb)])
    else:
        A = B.blockop([[B.JopDense(g.random((64, 48))) for _ in range(2)] for _ in range(2)])
    x = B.rand(B.domain(A), seed=1)
    out0 = B.rand(B.range_(A), seed=2)
    sa, so = B.solvers._S(1.7), B.solvers._S(0.3)
    for op in (A, B.adjoint(A)):
        xin, o0 = (x, out0) if op is A else (out0, x)
        got = o0.copy()
        B.solvers._apply_axpby(got, op, xin, sa, 0.0, L.COEF_INV, so, 0.0, L.COEF_NEG)
        ref = o0.copy()
        tmp = op * xin
        B.solvers._axpby(ref, sa, 0.0, L.COEF_INV, tmp, so, 0.0, L.COEF_NEG, ref)
        assert_bits(got.to_host(), ref.to_host())
        # jets_apply_axpby_norm: the same vector, and its norm from the store epilogue's partials (no norm pass for
        # the fused plan; a separate pass for the staged one) -- 1e-12 against numpy, and run-to-run identical
        nrm = B.solvers._S(0.0)
        vals = []
        for _ in range(3):
            got2 = o0.copy()
            B.solvers._apply_axpby_norm(got2, op, xin, sa, 0.0, L.COEF_INV, so, 0.0, L.COEF_NEG, nrm)
            vals.append(nrm.get())
            assert_bits(got2.to_host(), ref.to_host())
        want = np.linalg.norm(ref.to_host().astype(np.float64))
        assert abs(vals[0] - want) <= 1e-12 * want
        assert vals[0] == vals[1] == vals[2]


@pytest.mark.parametrize("T,n", [(np.float64, 100_003), (np.float32, 4096), (np.float32, 7)])
def test_axpby_pair_matches_two_updates(D, T, n):
    """jets_axpby_pair_dev: LSQR's x += t1 w ; w = v/alpha - t2 w in one pass, outputs aliasing inputs -- bit-identical
    to the two jets_axpby_dev calls in that order."""
    B = D.B
    L = B.solvers.L
    sp = B.JetSpace(T, n)
    x, w, v = B.rand(sp, seed=1), B.rand(sp, seed=2), B.rand(sp, seed=3)
    t1, t2, al = B.solvers._S(0.37), B.solvers._S(1.9), B.solvers._S(2.5)
    xr, wr = x.copy(), w.copy()
    B.solvers._axpby(xr, None, 1.0, 0, xr, t1, 0.0, 0, wr)
    B.solvers._axpby(wr, al, 0.0, L.COEF_INV, v, t2, 0.0, L.COEF_NEG, wr)
    B.solvers._axpby_pair(x, None, 1.0, 0, x, t1, 0.0, 0, w, w, al, 0.0, L.COEF_INV, v, t2, 0.0, L.COEF_NEG, w)
    assert_bits(x.to_host(), xr.to_host())
    assert_bits(w.to_host(), wr.to_host())


def test_vectorized_operator(O, D):
    """test/runtests.jl:797-838: B = vec(A) applied to vec(x) equals A*x, for a matrix-shaped space and
    for a block operator; vec(A') maps back to the domain."""
    T = np.float64
    g = np.random.default_rng(31)
    w = np.asfortranarray(g.random((10, 11)))
    xv = g.random((10, 11))

    def scn(K):
        A = K.JopDiagonal(w) if K.name == "oracle" else K.B.JopDiagonal(K.B.to_device(w, K.JetSpace(T, 10, 11)))
        x = K.arr(xv, K.domain(A))
        vec = K.J.vec if K.name == "oracle" else K.B.vec
        Bv = vec(A)
        assert tuple(K.domain(Bv).size()) == (110,) and tuple(K.domain(A).size()) == (10, 11)
        d = A * x
        _d = Bv * vec(x)
        A2 = K.blockop([[A], [A]])
        x2 = K.arr(xv, K.domain(A2))
        d2 = A2 * x2
        _d2 = vec(A2) * vec(x2)
        a2 = vec(K.adjoint(A2)) * d2
        return [np.asarray(K.host(v)).reshape(-1, order="F") for v in (d, _d, d2, _d2, a2)]
    o, dv = scn(O), scn(D)
    assert_bits(dv[0], dv[1])
    assert_bits(dv[2], dv[3])
    for a, b in zip(o, dv):
        assert_bits(a, b)


def test_copy_gives_an_independent_linearization_point(D):  # copy(A, false) src/Jets.jl:230-233; jacobian :374
    """B.copy(F) is a new jet: point! on the copy does not move the original's linearization point
    (what deepcopy(jet.s) guarantees inside Jets' own jacobian), while the state buffers are shared."""
    B = D.B
    g = np.random.default_rng(41)
    n = 2000
    w, m1, m2, dm = g.random(n), g.random(n), g.random(n), g.random(n)
    F = B.compose(B.JopDiagonal(w), B.JopPointwise(np.float64, n, "square"))
    J1 = B.jacobian_(F, B.to_device(m1))          # shares F's jet, linearized at m1
    G = B.copy(F)
    J2 = B.jacobian_(G, B.to_device(m2))          # the copy, linearized at m2
    assert np.array_equal((J1 * B.to_device(dm)).to_host(), w * ((2.0 * m1) * dm))
    assert np.array_equal((J2 * B.to_device(dm)).to_host(), w * ((2.0 * m2) * dm))
    assert np.array_equal((J1 * B.to_device(dm)).to_host(), w * ((2.0 * m1) * dm))   # still at m1
    assert np.array_equal((G * B.to_device(m1)).to_host(), (F * B.to_device(m1)).to_host())
    At = B.copy(B.JopDiagonal(w).T)
    assert np.array_equal((At * B.to_device(dm)).to_host(), w * dm)


def test_captured_graph_keeps_its_plans_alive(D):
    """A CUDA graph captured through jets_graph_begin/_end holds raw pointers into the plan tables of the applies it
    recorded.  point! replaces the operator's cached plan; the executable graph pins the one it captured, so a replay
    still computes -- with the linearization point it was captured with (the caller keeps that vector alive)."""
    import ctypes as C
    B = D.B
    g = np.random.default_rng(91)
    n = 20_000
    w, mo1, mo2, m = (g.random(n) for _ in range(4))
    F = B.JopDiagonal(w) @ B.JopPointwise(np.float64, n, "square")
    mo1_d, mo2_d = B.to_device(mo1), B.to_device(mo2)
    A = B.jacobian_(F, mo1_d)
    md, dd = B.to_device(m), B.zeros(B.range_(A))
    B.mul_(dd, A, md)                                   # builds the plan outside the capture
    ref1 = dd.to_host().copy()
    assert np.array_equal(ref1, w * (2 * mo1 * m))
    B.check(B.lib.jets_graph_begin())
    try:
        B.mul_(dd, A, md)
    finally:
        gh = C.c_void_p()
        B.check(B.lib.jets_graph_end(C.byref(gh)))
    try:
        B.jacobian_(F, mo2_d)                           # new point on the SHARED jet (:364-366): A's cached plan is stale ...
        B.mul_(dd, A, md)                               # ... and is rebuilt here (the old one is dropped from the cache)
        assert np.array_equal(dd.to_host(), w * (2 * mo2 * m))
        dd.fill_(0.0)
        for _ in range(3):
            B.check(B.lib.jets_graph_launch(gh))        # ... while the graph replays the plan it captured
        B.sync()
        assert np.array_equal(dd.to_host(), ref1)
    finally:
        B.check(B.lib.jets_graph_destroy(gh))


def test_jet_accessor(O, D):
    """jet(A) (src/Jets.jl:236-238, exported): F and its JopLn view share ONE jet, the adjoint's jet is its parent's,
    jacobian(F, mo) makes a new one; the accessors that take a Jop take the Jet (domain / range / shape / state /
    point / point!).  A Jet cannot be built from host closures on the device path."""
    B, J = D.B, O.J
    g = np.random.default_rng(92)
    n = 4000
    w, mo1, mo2, m = (g.random(n) for _ in range(4))
    F = B.JopDiagonal(w) @ B.JopPointwise(np.float64, n, "square")
    A = B.jacobian_(F, B.to_device(mo1))
    assert B.jet(F) is B.jet(A) and B.jet(B.adjoint(A)) is B.jet(A) and B.jet(B.jet(A)) is B.jet(A)
    j = B.jet(A)
    assert isinstance(j, B.Jet) and B.domain(j) == B.domain(A) and B.range_(j) == B.range_(A) and B.shape(j) == B.shape(A)
    assert B.state(j) is B.state(A) and np.array_equal(B.point(j).to_host(), mo1)
    Jn = B.jacobian(F, B.to_device(mo1))
    assert B.jet(Jn) is not B.jet(F)                       # jacobian copies the jet (:374)
    B.point_(j, B.to_device(mo2))                          # point!(jet(F), mo): every operator on this jet sees it
    assert np.array_equal((A * B.to_device(m)).to_host(), w * (2 * mo2 * m))
    assert np.array_equal((Jn * B.to_device(m)).to_host(), w * (2 * mo1 * m))
    # the oracle's jets behave the same way
    Fo = J.JopDiagonal(w) @ J.JopPointwise(np.float64, n, "square")
    Ao = J.jacobian_(Fo, mo1.copy())
    assert J.jet(Fo) is J.jet(Ao) and J.jet(J.adjoint(Ao)) is J.jet(Ao) and J.jet(J.jacobian(Fo, mo1.copy())) is not J.jet(Fo)
    with pytest.raises(B.JetsError):
        B.Jet(dom=B.JetSpace(np.float64, n), rng=B.JetSpace(np.float64, n), f=lambda d, m: d)
