"""GPU parity tests for the input-tile-cache ("bundle") engine of the fused block apply
(csrc/kernels_fused_bundle.cu): rows that share an input block (`_m = getblock(m, jblock)`,
src/Jets.jl:1019; `_d`, :1044) fetch its tiles from HBM/L2 once and reuse them from shared memory.

Every scenario is evaluated by the numpy oracle and by the device on the same seeded inputs and
must be BIT-IDENTICAL (elementwise/stencil arithmetic, one rounding per op, fixed left-to-right row
sums as :1024/:1049), and the three device engines (cache / streaming TMA / guarded loads) must
agree with each other bit for bit.  The last test re-runs this file in sub-processes with tiny
rings (JETS_B200_BUNDLE_NX/NS/BMAX) so that ring recycling, group splitting and bundle splitting
are exercised at sizes the oracle finishes in seconds.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from backends import OracleBackend, DeviceBackend

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_KEEP = {}


@pytest.fixture(scope="module")
def O():
    return OracleBackend()


@pytest.fixture(scope="module")
def D():
    return DeviceBackend()


def assert_bits(a, b):
    a, b = np.atleast_1d(np.asarray(a)), np.atleast_1d(np.asarray(b))
    assert a.dtype == b.dtype and a.shape == b.shape, (a.dtype, b.dtype, a.shape, b.shape)
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), \
        f"not bit-identical: max abs diff {np.max(np.abs(a.astype(np.float64) - b.astype(np.float64))):.3e}"


def both_ways(O, D, build, T, seed=0, expect_cache=True):
    """Builds the operator with both backends, applies forward and adjoint, bit-compares against the
    oracle and across the device engines."""
    g = np.random.default_rng(seed)
    Ao, Ad = build(O), build(D)
    m = g.random(len(O.domain(Ao))).astype(T)
    d = g.random(len(O.range_(Ao))).astype(T)
    fo = O.host(Ao * O.arr(m, O.domain(Ao)))
    to = O.host(O.adjoint(Ao) * O.arr(d, O.range_(Ao)))
    B = D.B
    md, dd = D.arr(m, D.domain(Ad)), D.arr(d, D.range_(Ad))
    res = {}
    for eng in ("auto", "tma_nocache", "ldg"):
        B.set_fused_engine(eng)
        try:
            A = build(D)
            res[eng] = (D.host(A * md), D.host(D.adjoint(A) * dd), B.plan_info(A))
        finally:
            B.set_fused_engine("auto")
    for eng, (f, t, info) in res.items():
        assert_bits(f, fo)
        assert_bits(t, to)
    assert res["auto"][2]["launches"] == 1
    if expect_cache:   # 16-byte aligned blocks: the TMA engine with the input-tile cache
        assert res["auto"][2]["engines"] == ["tma"], res["auto"][2]
        # (with forced ring sizes the planner may legitimately fall back to the streaming kernel)
        assert res["auto"][2]["input_cache"] or os.environ.get("JETS_B200_BUNDLE_SUBPROCESS"), res["auto"][2]
    else:              # block starts off 16-byte boundaries: guarded-load engine
        assert res["auto"][2]["engines"] == ["ldg"], res["auto"][2]
    assert not res["tma_nocache"][2]["input_cache"]
    lhs, rhs = D.dot_product_test(Ad, md, dd)
    tol = 1e-12 if T == np.float64 else 1e-5
    assert abs(lhs - rhs) <= tol * abs(lhs + rhs)


@pytest.mark.parametrize("T", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 36, 4096, 50_004])
def test_blockdiag_4x4_config1_shape(O, D, T, n):
    """Config 1 shape: every x block is read by all four rows -> one bundle of four rows."""
    g = np.random.default_rng(1)
    W = [[g.random(n).astype(T) for _ in range(4)] for _ in range(4)]
    both_ways(O, D, lambda K: K.blockop([[K.JopDiagonal(W[r][c]) for c in range(4)] for r in range(4)]), T, seed=n,
              expect_cache=(n * np.dtype(T).itemsize) % 16 == 0)


@pytest.mark.parametrize("T", [np.float32, np.float64])
@pytest.mark.parametrize("nb,n", [(2, 8), (12, 4100), (40, 20_004)])
def test_block_tridiagonal_config5_shape(O, D, T, nb, n):
    """Config 5 shape: the input tile window slides down the block rows (3 live tiles)."""
    g = np.random.default_rng(2)
    W = [g.random(n).astype(T) for _ in range(nb)]

    def build(K):
        sp = K.JetSpace(T, n)
        return K.blockop([[K.JopDiagonal(W[r]) if r == c else K.JopStencil(T, n, "fdiff") if c == r + 1 else
                           K.JopStencil(T, n, "lap") if c == r - 1 else K.JopZeroBlock(sp, sp)
                           for c in range(nb)] for r in range(nb)])
    both_ways(O, D, build, T, seed=nb)


@pytest.mark.parametrize("shape", [(1, 23), (23, 1), (6, 6), (3, 19)])
def test_many_inputs_and_many_rows(O, D, shape):
    """1 x C rows allocate more input tiles than the ring holds (recycling inside one row);
    R x 1 columns share one input across all rows; 6x6 needs more live tiles than fit."""
    T = np.float32
    R, C_ = shape
    n = 9004
    g = np.random.default_rng(3)
    W = [[g.random(n).astype(T) for _ in range(C_)] for _ in range(R)]

    def build(K):
        return K.blockop([[K.JopDiagonal(W[r][c]) if (r + c) % 3 else 0.5 * K.JopStencil(T, n, "lap")
                           for c in range(C_)] for r in range(R)])
    both_ways(O, D, build, T, seed=R * 100 + C_)


@pytest.mark.parametrize("T", [np.float32, np.float64])
@pytest.mark.parametrize("n", [4, 4100, 33_000])
def test_weighted_stencil_blocks(O, D, T, n):
    """w .* S(x) and S(w .* x) blocks (weighted derivatives and their adjoints) have straight-line kernel
    paths (PAT_*_DIAG / PAT_DIAG_*): bit-identical to the oracle's stage-by-stage evaluation, block edges
    included (the stencil needs the halo element of BOTH streams)."""
    g = np.random.default_rng(31)
    nb = 4
    W = [[g.random(n).astype(T) - 0.5 for _ in range(nb)] for _ in range(nb)]

    def build(K):
        def blk(r, c):
            S = K.JopStencil(T, n, ("fdiff", "lap")[(r + c) % 2])
            Dg = K.JopDiagonal(W[r][c])
            k = (2 * r + c) % 4
            return Dg @ S if k == 0 else S @ Dg if k == 1 else K.adjoint(Dg @ S) if k == 2 else Dg
        return K.blockop([[blk(r, c) for c in range(nb)] for r in range(nb)])
    both_ways(O, D, build, T, seed=n)


def test_ragged_rows_split_bundles(O, D):
    """Rows of different length cannot share a tile position: bundles close at every length change;
    zero blocks leave rows without terms (zero-filled) in the middle of the operator."""
    T = np.float64
    lens = [3000, 3000, 16, 5000, 5000, 5000, 2]
    g = np.random.default_rng(4)
    nb = len(lens)
    W = {(r, c): g.random(lens[r]) for r in range(nb) for c in range(nb) if lens[r] == lens[c]}

    def build(K):
        def blk(r, c):
            if lens[r] != lens[c] or r == 2:
                return K.JopZeroBlock(K.JetSpace(T, lens[c]), K.JetSpace(T, lens[r]))
            if r == c:
                return 2.5 * K.JopStencil(T, lens[r], "fdiff")
            if c == r + 1:   # (c*S')' = S o c: scale-then-stencil chains in the adjoint
                return 0.75 * K.adjoint(K.JopStencil(T, lens[r], "fdiff"))
            return K.JopDiagonal(W[(r, c)])
        return K.blockop([[blk(r, c) for c in range(nb)] for r in range(nb)])
    both_ways(O, D, build, T, seed=5)


@pytest.mark.parametrize("T", [np.float32, np.float64])
def test_sum_shares_input_inside_a_row(O, D, T):
    """B - c*S (config 4): both terms of a row read the same block -> one tile, two uses."""
    nb, n = 5, 30_012
    g = np.random.default_rng(6)
    W = [1.0 + g.random(n).astype(T) for _ in range(nb)]

    def build(K):
        sp = K.JetSpace(T, n)
        Bd = K.blockop([[K.JopDiagonal(W[i]) if i == j else K.JopZeroBlock(sp, sp) for j in range(nb)] for i in range(nb)])
        Sd = K.blockop([[K.JopStencil(T, n, "lap") if i == j else K.JopZeroBlock(sp, sp) for j in range(nb)] for i in range(nb)])
        return Bd - 0.5 * Sd
    both_ways(O, D, build, T, seed=7)


@pytest.mark.parametrize("T", [np.float32, np.float64])
def test_jacobian_chain_blocks(O, D, T):
    """Blocks of D ∘ S ∘ J(x^2) sharing inputs AND linearization points across rows."""
    nb, n = 3, 12_348
    g = np.random.default_rng(8)
    W = [[g.random(n).astype(T) for _ in range(nb)] for _ in range(nb)]
    mo = [g.random(n).astype(T) for _ in range(nb)]

    def build(K):
        sp = K.JetSpace(T, n)
        Jc = [K.jacobian(K.JopPointwise(T, n, "square"), K.arr(mo[c], sp)) for c in range(nb)]
        return K.blockop([[K.JopDiagonal(W[r][c]) @ K.JopStencil(T, n, "fdiff") @ Jc[c] for c in range(nb)]
                          for r in range(nb)])
    both_ways(O, D, build, T, seed=9)


def test_accumulate_into_dirty_output(O, D):
    """Quirk Q1 (src/Jets.jl:1001,1024): mul!(d, A, m) with ncol>1 adds into d."""
    T = np.float64
    n = 7000
    g = np.random.default_rng(10)
    W = [[g.random(n) for _ in range(3)] for _ in range(3)]
    m, d0 = g.random(3 * n), g.random(3 * n)

    def scn(K):
        A = K.blockop([[K.JopDiagonal(W[r][c]) for c in range(3)] for r in range(3)])
        d = K.arr(d0, K.range_(A))
        if K.name == "device":
            K.B.mul_(d, A, K.arr(m, K.domain(A)), accumulate=True)
        else:
            K.mul_(d, A, K.arr(m, K.domain(A)))
        return K.host(d)
    assert_bits(scn(D), scn(O))


def test_large_rows_many_units(D):
    """Several tiles per row and more units than SMs: checked through the size-independent
    properties (dot-product test, linearity, engine agreement) -- no oracle at this size."""
    B = D.B
    T = np.float32
    n = 3_000_016
    sp = B.JetSpace(T, n)
    nb = 6
    Wd = B.rand(B.JetBSpace([sp] * nb), seed=11)
    A = B.blockop([[B.JopDiagonal(B.getblock(Wd, r + 1)) if r == c else B.JopStencil(T, n, "fdiff") if c == r + 1 else
                    B.JopStencil(T, n, "lap") if c == r - 1 else B.JopZeroBlock(sp, sp) for c in range(nb)]
                   for r in range(nb)])
    m, d = B.rand(B.domain(A), seed=12), B.rand(B.range_(A), seed=13)
    lhs, rhs = B.dot_product_test(A, m, d)
    assert abs(lhs - rhs) <= 1e-5 * abs(lhs + rhs)
    f1, t1 = (A * m).to_host(), (A.T * d).to_host()
    assert B.plan_info(A)["input_cache"]
    B.set_fused_engine("tma_nocache")
    try:
        A2 = B.blockop([[B.JopDiagonal(B.getblock(Wd, r + 1)) if r == c else B.JopStencil(T, n, "fdiff") if c == r + 1 else
                         B.JopStencil(T, n, "lap") if c == r - 1 else B.JopZeroBlock(sp, sp) for c in range(nb)]
                        for r in range(nb)])
        f2, t2 = (A2 * m).to_host(), (A2.T * d).to_host()
    finally:
        B.set_fused_engine("auto")
    assert_bits(f1, f2)
    assert_bits(t1, t2)


@pytest.mark.parametrize("nchunks", [1, 3, 12])
def test_chunked_host_pipeline_matches_monolithic_apply(D, nchunks):
    """jets_dist_apply_normal_host (host m -> A -> A' -> host m', chunked over block rows on three streams
    inside the library; here one rank, no neighbours) must reproduce the monolithic device applies bit for
    bit, step after step -- through jets_dist_apply AND through plain jets_apply on the same block rows."""
    import torch
    B = D.B
    T = np.float32
    nblk, n = 12, 40_000
    part = B.dist.RowPartition(nblk, 1, 0, halo=1)
    sp = B.JetSpace(T, n)
    W = B.rand(B.JetBSpace([sp] * nblk), seed=21)
    Z = B.JopZeroBlock(sp, sp)

    def make_block(r, c):
        if r == c:
            return B.JopDiagonal(B.getblock(W, r + 1))
        return B.JopStencil(T, n, "fdiff") if c == r + 1 else B.JopStencil(T, n, "lap")
    A = B.dist.build_local_operator(B, part, make_block, lambda: Z)     # nblk x (nblk + 2), zero halo columns
    op = B.dist.DistOp(B, A, halo=1)
    own = B.JetBSpace([sp] * nblk)
    g = np.random.default_rng(22)
    h_in = torch.empty(nblk * n, dtype=torch.float32, pin_memory=True)
    h_out = torch.empty(nblk * n, dtype=torch.float32, pin_memory=True)
    try:
        for it in range(3):
            m = g.random(nblk * n).astype(T)
            h_in.copy_(torch.from_numpy(m))
            h_out.zero_()
            op.normal_host(h_out.data_ptr(), h_in.data_ptr(), nchunks)
            op.join()
            B.sync()
            got = h_out.numpy().copy()
            assert op.info(3) == min(nchunks, nblk)
            # one jets_dist_apply per direction on device-resident shards
            x, d2, m2 = B.to_device(m, own), B.zeros(own), B.zeros(own)
            op.forward(d2, x)
            op.adjoint(m2, d2)
            assert_bits(got, m2.to_host())
            # and the plain operator path over the halo-extended vectors (halo blocks zero, unread)
            x_ext, m_ext, d3 = B.zeros(B.domain(A)), B.zeros(B.domain(A)), B.zeros(B.range_(A))
            B.reshape(x_ext, B.JetSpace(T, (nblk + 2) * n)).from_host(np.concatenate([np.zeros(n, T), m, np.zeros(n, T)]))
            B.mul_(d3, A, x_ext)
            B.mul_(m_ext, B.adjoint(A), d3)
            assert_bits(got, m_ext.to_host()[n:-n])
    finally:
        B.sync()
        op.close()


@pytest.mark.parametrize("env", [
    {"JETS_B200_BUNDLE_NX": "4", "JETS_B200_BUNDLE_NS": "2"},
    {"JETS_B200_BUNDLE_NX": "1", "JETS_B200_BUNDLE_NS": "1"},
    {"JETS_B200_BUNDLE_NX": "5", "JETS_B200_BUNDLE_NS": "3", "JETS_B200_BUNDLE_BMAX": "3"},
    {"JETS_B200_FAST_VARIANT": "0"},
    {"JETS_B200_FAST_VARIANT": "1", "JETS_B200_BUNDLE_NX": "3"},
    {"JETS_B200_NO_PDL": "1"},
    {"JETS_B200_TAIL_MIN_UNITS": "1"},                                  # fine-grained tail sub-bundles at test sizes
    {"JETS_B200_TAIL_MIN_UNITS": "1", "JETS_B200_BUNDLE_NX": "4", "JETS_B200_BUNDLE_NS": "2"},
    {"JETS_B200_NO_PRE_STATE": "1", "JETS_B200_NO_FIRST_STATIC": "1"},
])
def test_tiny_rings_and_other_tile_shapes(env):
    """Re-runs the scenarios above with ring sizes that force recycling waits, group splits and
    bundle splits (the options are read once, at jets_init)."""
    if os.environ.get("JETS_B200_BUNDLE_SUBPROCESS"):
        pytest.skip("already inside the sub-process run")
    e = dict(os.environ)
    e.update(env)
    e["JETS_B200_BUNDLE_SUBPROCESS"] = "1"
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-m", "gpu",
                        "-k", "not tiny_rings"], cwd=ROOT, env=e, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
