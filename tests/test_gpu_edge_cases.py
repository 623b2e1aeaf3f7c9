"""GPU edge cases of the operator-application path: empty and one-element blocks, ragged block
lengths that straddle tile boundaries, block rows made only of zero blocks, operators applied to
views, empty restrictions -- each against the numpy oracle on the same seeded inputs (bit-exact: these
are elementwise / stencil / data-movement paths).  The reference's own tests use tiny shapes
(test/runtests.jl:512-787: blocks of 2, 4 and 6 elements); the device engines tile at 8-16 KB, so the
interesting sizes here sit around the tile and vector boundaries."""
import numpy as np
import pytest

from backends import OracleBackend, DeviceBackend

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def O():
    return OracleBackend()


@pytest.fixture(scope="module")
def D():
    return DeviceBackend()


def bits(a, b):
    a, b = np.atleast_1d(np.asarray(a)), np.atleast_1d(np.asarray(b))
    return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))


def both(O, D, build, lens_in, lens_out, T, seed):
    g = np.random.default_rng(seed)
    m = g.random(sum(lens_in)).astype(T)
    d = g.random(sum(lens_out)).astype(T)
    res = []
    for K in (O, D):
        A = build(K)
        R, Dm = K.range_(A), K.domain(A)
        f = K.host(K.mul_(K.zeros(R), A, K.arr(m, Dm)))
        t = K.host(K.mul_(K.zeros(Dm), K.adjoint(A), K.arr(d, R)))
        res.append((np.asarray(f).reshape(-1, order="F"), np.asarray(t).reshape(-1, order="F")))
    return res


@pytest.mark.parametrize("T", [np.float32, np.float64])
@pytest.mark.parametrize("lens", [(1, 1, 1), (1, 2047, 2049, 3), (4096, 4097, 1, 8191), (5, 0, 7), (0, 0, 3)])
def test_block_diagonal_ragged_and_empty_blocks(O, D, T, lens):
    """Block-diagonal operator of diagonal + stencil blocks over ragged block lengths, including empty
    (length-0) blocks: block offsets (src/Jets.jl:742-748) and every tile edge are exercised."""
    ws = [np.random.default_rng(100 + i).random(n).astype(T) for i, n in enumerate(lens)]

    def build(K):
        rows = []
        for i, n in enumerate(lens):
            row = []
            for j, k in enumerate(lens):
                if i == j:
                    op = K.JopDiagonal(K.arr(ws[i], K.JetSpace(T, n)) if K.name == "device" else ws[i])
                    if n >= 1:
                        op = op - K.JopStencil(T, n, "lap") if n >= 1 else op
                    row.append(op)
                else:
                    row.append(K.JopZeroBlock(K.JetSpace(T, k), K.JetSpace(T, n)))
            rows.append(row)
        return K.blockop(rows)

    (fo, to), (fd, td) = both(O, D, build, lens, lens, T, 1)
    assert bits(fo, fd) and bits(to, td)


@pytest.mark.parametrize("T", [np.float32, np.float64])
def test_rows_and_columns_of_zero_blocks(O, D, T):
    """A block row (and a block column) made only of JopZeroBlock: the forward zero-fills that range
    block through `A*m` (:399 passes zeros(range(A))), the adjoint zero-fills the domain block (:1041)."""
    n = 1500
    w = np.random.default_rng(5).random(n).astype(T)

    def build(K):
        Z = lambda: K.JopZeroBlock(K.JetSpace(T, n), K.JetSpace(T, n))
        Dg = K.JopDiagonal(K.arr(w, K.JetSpace(T, n)) if K.name == "device" else w)
        S = K.JopStencil(T, n, "fdiff")
        return K.blockop([[Dg, Z(), S], [Z(), Z(), Z()], [S, Z(), Dg]])

    (fo, to), (fd, td) = both(O, D, build, (n, n, n), (n, n, n), T, 2)
    assert bits(fo, fd) and bits(to, td)
    assert not fd[n:2 * n].any() and not td[n:2 * n].any()


def test_one_by_one_and_single_row_single_column_blocks(O, D):  # runtests.jl:724-757
    T, n = np.float64, 777
    w = [np.random.default_rng(20 + i).random(n) for i in range(3)]

    def row(K):
        return K.blockop([[K.JopDiagonal(K.arr(v, K.JetSpace(T, n)) if K.name == "device" else v) for v in w]])

    def col(K):
        return K.blockop([[K.JopDiagonal(K.arr(v, K.JetSpace(T, n)) if K.name == "device" else v)] for v in w])

    def one(K):
        return K.blockop([[K.JopDiagonal(K.arr(w[0], K.JetSpace(T, n)) if K.name == "device" else w[0])]])

    for build, li, lo in ((row, (n,) * 3, (n,)), (col, (n,), (n,) * 3), (one, (n,), (n,))):
        (fo, to), (fd, td) = both(O, D, build, li, lo, T, 3)
        assert bits(fo, fd) and bits(to, td)


def test_empty_restriction_and_full_permutation(D):
    B = D.B
    n = 1000
    g = np.random.default_rng(9)
    m = g.random(n)
    R0 = B.JopRestriction(np.float64, n, [])
    assert (R0 * B.to_device(m)).to_host().size == 0
    back = (R0.T * B.zeros(B.JetSpace(np.float64, 0))).to_host()
    assert back.shape == (n,) and not back.any()
    perm = g.permutation(n) + 1
    P = B.JopRestriction(np.float64, n, perm)
    assert np.array_equal((P * B.to_device(m)).to_host(), m[perm - 1])
    assert np.array_equal((P.T * (P * B.to_device(m))).to_host(), m)      # a permutation is orthogonal


def test_vector_ops_on_tiny_and_unaligned_views(D):
    """dot/norm/lincomb/fill on 0-, 1- and odd-length blocks and on block views that start at
    unaligned offsets inside the flat buffer."""
    B = D.B
    g = np.random.default_rng(10)
    for T in (np.float32, np.float64, np.complex64):
        R = B.JetBSpace([B.JetSpace(T, 1), B.JetSpace(T, 0), B.JetSpace(T, 3), B.JetSpace(T, 1025)])
        xh = g.random(len(R)).astype(T)
        yh = g.random(len(R)).astype(T)
        x, y = B.to_device(xh, R), B.to_device(yh, R)
        tol = 1e-5 if np.dtype(T).itemsize <= 8 and np.dtype(T) != np.float64 else 1e-12
        for i in (1, 2, 3, 4):
            xb, yb = B.getblock(x, i), B.getblock(y, i)
            a, b = xh[R.indices[i - 1][0] - 1:R.indices[i - 1][1]], yh[R.indices[i - 1][0] - 1:R.indices[i - 1][1]]
            assert abs(complex(B.dot(xb, yb)) - np.vdot(a.astype(np.complex128), b.astype(np.complex128))) <= tol * max(1.0, abs(np.vdot(a, b)))
            assert abs(float(B.norm(xb)) - np.linalg.norm(a.astype(np.complex128))) <= tol * max(1.0, np.linalg.norm(a))
            if len(xb):
                s = (xb * 2.0 + yb).to_host().reshape(-1)
                assert np.allclose(s, 2 * a + b, rtol=10 * tol)
        B.getblock(x, 3).fill_(7.0)
        h = x.to_host()
        assert np.all(h[1:4] == 7) and h[0] == xh[0] and h[4] == xh[4]
