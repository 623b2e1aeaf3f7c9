"""The C restatement (CPU baseline) must agree bit for bit with the numpy oracle in every mode."""
import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import jets_oracle as J


@pytest.mark.parametrize("T", [np.float32, np.float64])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_block_tridiagonal_matches_numpy_oracle(T, mode):
    g = np.random.default_rng(0)
    nb, n = 5, 4099
    W = [g.random(n).astype(T) for _ in range(nb)]
    sp = J.JetSpace(T, n)

    def leaf(r, c):
        if r == c:
            return ("diag", W[r]), J.JopDiagonal(W[r])
        if c == r + 1:
            return ("fdiff", None), J.JopStencil(T, n, "fdiff")
        if c == r - 1:
            return ("lap", None), J.JopStencil(T, n, "lap")
        return ("zero", None), J.JopZeroBlock(sp, sp)
    leaves = [[leaf(r, c)[0] for c in range(nb)] for r in range(nb)]
    A = J.blockop([[leaf(r, c)[1] for c in range(nb)] for r in range(nb)])
    Ac = CO.BlockOp(leaves, [n] * nb, [n] * nb, T)
    m = g.random(nb * n).astype(T)
    d = g.random(nb * n).astype(T)
    f_ref = J.to_array(A * J.reshape(m.copy(), J.domain(A)))
    t_ref = J.to_array(A.T * J.reshape(d.copy(), J.range_(A)))
    assert np.array_equal(Ac.apply(m, False, mode), f_ref)
    assert np.array_equal(Ac.apply(d, True, mode), t_ref)


@pytest.mark.parametrize("T", [np.float32, np.float64])
@pytest.mark.parametrize("mode", [0, 1])
def test_chain_matches_numpy_oracle(T, mode):
    g = np.random.default_rng(1)
    n = 10007
    w, mo, dm, dd = (g.random(n).astype(T) for _ in range(4))
    A = J.JopDiagonal(w) @ J.JopStencil(T, n, "fdiff") @ J.jacobian(J.JopPointwise(T, n, "square"), mo)
    assert np.array_equal(CO.chain_apply(w, mo, dm, False, mode), A * dm)
    assert np.array_equal(CO.chain_apply(w, mo, dd, True, mode), A.T * dd)
