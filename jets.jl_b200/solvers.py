"""LSQR-style caller of the hot path: every iteration is A*v, A'*u, two norms and
a handful of BlockArray broadcast updates -- exactly what IterativeSolvers.lsqr does with a Jets
operator (docs/src/index.md:235-246).  Written only with the public API of this package so that
it exercises mul!, adjoint mul!, norm and the broadcast updates the way a Jets user would.

``lsqr``         host scalars: two stream synchronisations per iteration (the norms).
``lsqr_graph``   device scalars + one CUDA graph per iteration: no host round trip at all.
"""
from __future__ import annotations

import ctypes as C

from . import _lib as L
from ._lib import check, lib
from . import core as J


def lsqr(A, b, iters=10, x0=None):
    """Golub-Kahan LSQR (Paige & Saunders 1982) without stopping tests: exactly ``iters``
    iterations.  Returns (x, history of (alpha, beta))."""
    x = J.zeros(J.domain(A)) if x0 is None else x0
    At = J.adjoint(A)
    u = b.copy()
    beta = float(J.norm(u))
    J.lincomb_(u, [(1.0 / beta, u)])
    v = At * u
    alpha = float(J.norm(v))
    J.lincomb_(v, [(1.0 / alpha, v)])
    w = v.copy()
    phibar, rhobar = beta, alpha
    tmp_u = J.zeros(J.range_(A))
    tmp_v = J.zeros(J.domain(A))
    hist = []
    for _ in range(iters):
        J.mul_(tmp_u, A, v)                                  # u = A v - alpha u
        J.lincomb_(u, [(1.0, tmp_u), (-alpha, u)])
        beta = float(J.norm(u))
        J.lincomb_(u, [(1.0 / beta, u)])
        J.mul_(tmp_v, At, u)                                 # v = A' u - beta v
        J.lincomb_(v, [(1.0, tmp_v), (-beta, v)])
        alpha = float(J.norm(v))
        J.lincomb_(v, [(1.0 / alpha, v)])
        rho = (rhobar * rhobar + beta * beta) ** 0.5
        c, s = rhobar / rho, beta / rho
        theta = s * alpha
        rhobar = -c * alpha
        phi = c * phibar
        phibar = s * phibar
        J.lincomb_(x, [(1.0, x), (phi / rho, w)])            # x += (phi/rho) w
        J.lincomb_(w, [(1.0, v), (-theta / rho, w)])         # w = v - (theta/rho) w
        hist.append((alpha, beta))
    return x, hist


def lsqr_dist(op, b_own, iters=10):
    """The same iteration on a block-row partitioned operator (``dist.DistOp``): every rank holds its shards of
    u (range) and v, w, x (domain); ``A*v`` / ``A'*u`` are one ``jets_dist_apply`` each (halo exchange inside the
    kernel launch), the norms are local sums of squares added over the ranks in rank order
    (``jets_dist_sum_scalar``: bit-stable, the same on every rank) -- north_star's "scalars with all-reduce".
    Returns (x_own, history of (alpha, beta)); the iterates equal ``lsqr`` on the whole operator up to the
    rounding of the norms."""
    B = op.B

    def gnorm(x):
        r = C.c_double()
        check(lib.jets_dot(x._h, x._h, C.byref(r)))          # local sum of squares in Float64
        t = C.c_double(r.value)
        if lib.jets_dist_size() > 1:
            check(lib.jets_dist_sum_scalar(C.byref(t)))
        return t.value ** 0.5
    x = J.zeros(op.own_space)
    u = b_own.copy()
    beta = gnorm(u)
    J.lincomb_(u, [(1.0 / beta, u)])
    v = J.zeros(op.own_space)
    op.register(v)                                           # forward applies read the neighbours' halo of v in place
    op.adjoint(v, u)
    alpha = gnorm(v)
    J.lincomb_(v, [(1.0 / alpha, v)])
    w = v.copy()
    phibar, rhobar = beta, alpha
    tmp_u = J.zeros(op.range_space)
    tmp_v = J.zeros(op.own_space)
    hist = []
    for _ in range(iters):
        op.forward(tmp_u, v)                                 # u = A v - alpha u
        J.lincomb_(u, [(1.0, tmp_u), (-alpha, u)])
        beta = gnorm(u)
        J.lincomb_(u, [(1.0 / beta, u)])
        op.adjoint(tmp_v, u)                                 # v = A' u - beta v
        J.lincomb_(v, [(1.0, tmp_v), (-beta, v)])
        alpha = gnorm(v)
        J.lincomb_(v, [(1.0 / alpha, v)])
        rho = (rhobar * rhobar + beta * beta) ** 0.5
        c, s = rhobar / rho, beta / rho
        theta = s * alpha
        rhobar = -c * alpha
        phi = c * phibar
        phibar = s * phibar
        J.lincomb_(x, [(1.0, x), (phi / rho, w)])
        J.lincomb_(w, [(1.0, v), (-theta / rho, w)])
        hist.append((alpha, beta))
    return x, hist


class _S:
    """Device scalar (jets_scalar)."""

    def __init__(self, v=0.0):
        L.ensure_init()
        self.h = C.c_void_p()
        check(lib.jets_scalar_create(C.byref(self.h)))
        if v:
            self.set(v)

    def set(self, v):
        check(lib.jets_scalar_set(self.h, float(v)))

    def get(self):
        r = C.c_double()
        check(lib.jets_scalar_get(self.h, C.byref(r)))
        return r.value

    def __del__(self):
        try:
            lib.jets_scalar_destroy(self.h)
        except Exception:
            pass


def _sop(out, op, a, b=None):
    check(lib.jets_scalar_op(out.h, op.encode(), a.h, b.h if b is not None else None))


def _axpby(out, sa, ca, af, x, sb=None, cb=0.0, bf=0, y=None):
    check(lib.jets_axpby_dev(out._h, sa.h if sa is not None else None, ca, af, x._h,
                             sb.h if sb is not None else None, cb, bf, y._h if y is not None else None))


class LsqrGraph:
    """Golub-Kahan LSQR with every scalar resident on the device; one iteration is captured once in
    a CUDA graph and replayed (no host synchronisation inside the loop).  ``LsqrGraph(A, b)`` runs
    the start-up and the first iteration and captures the graph; ``run(k)`` replays k further
    iterations asynchronously on the library stream; ``result()`` synchronises."""

    def __init__(self, A, b):
        self.x = x = J.zeros(J.domain(A))
        At = J.adjoint(A)
        u = b.copy()
        beta, alpha, rho, rhobar, phibar, phi, theta, c, s, t1, t2 = (_S() for _ in range(11))
        self.alpha, self.beta = alpha, beta
        check(lib.jets_norm_dev(u._h, 2.0, beta.h))
        _axpby(u, beta, 0.0, L.COEF_INV, u)
        v = At * u
        check(lib.jets_norm_dev(v._h, 2.0, alpha.h))
        _axpby(v, alpha, 0.0, L.COEF_INV, v)
        w = v.copy()
        _sop(phibar, "+", beta)      # phibar = beta  (x + 0)
        _sop(rhobar, "+", alpha)
        tmp_u = J.zeros(J.range_(A))
        tmp_v = J.zeros(J.domain(A))
        self._keep = (A, At, u, v, w, tmp_u, tmp_v, rho, rhobar, phibar, phi, theta, c, s, t1, t2)

        def body():
            J.mul_(tmp_u, A, v)
            _axpby(u, None, 1.0, 0, tmp_u, alpha, 0.0, L.COEF_NEG, u)        # u = A v - alpha u
            check(lib.jets_norm_dev(u._h, 2.0, beta.h))
            _axpby(u, beta, 0.0, L.COEF_INV, u)
            J.mul_(tmp_v, At, u)
            _axpby(v, None, 1.0, 0, tmp_v, beta, 0.0, L.COEF_NEG, v)         # v = A' u - beta v
            check(lib.jets_norm_dev(v._h, 2.0, alpha.h))
            _axpby(v, alpha, 0.0, L.COEF_INV, v)
            _sprog([(rho, "h", rhobar, beta), (c, "/", rhobar, rho), (s, "/", beta, rho), (theta, "*", s, alpha),
                    (t1, "*", c, alpha), (rhobar, "n", t1, None), (phi, "*", c, phibar), (phibar, "*", s, phibar),
                    (t1, "/", phi, rho), (t2, "/", theta, rho)])      # the scalar recurrences: one launch
            _axpby(x, None, 1.0, 0, x, t1, 0.0, 0, w)                        # x += (phi/rho) w
            _axpby(w, None, 1.0, 0, v, t2, 0.0, L.COEF_NEG, w)               # w = v - (theta/rho) w

        body()  # iteration 1; builds every plan outside the capture
        check(lib.jets_graph_begin())
        try:
            body()  # captured, not executed
        finally:
            g = C.c_void_p()
            check(lib.jets_graph_end(C.byref(g)))
        self._g = g

    def run(self, iters):
        for _ in range(iters):
            check(lib.jets_graph_launch(self._g))

    def result(self):
        J.sync()
        return self.x, (self.alpha.get(), self.beta.get())

    def __del__(self):
        try:
            if self._g:
                lib.jets_graph_destroy(self._g)
        except Exception:
            pass


def _apply_axpby(out, A, x, sa, ca, af, so, co, of):
    """out = cA*(A x) + cO*out with device-resident coefficients (jets_apply_axpby)."""
    check(lib.jets_apply_axpby(A._h.h, A._mode, out._h, x._h, sa.h if sa is not None else None, ca, af,
                               so.h if so is not None else None, co, of))


def _apply_axpby_norm(out, A, x, sa, ca, af, so, co, of, nrm):
    """The same, and nrm = norm(out) from the store epilogue's partials: no norm pass (jets_apply_axpby_norm)."""
    check(lib.jets_apply_axpby_norm(A._h.h, A._mode, out._h, x._h, sa.h if sa is not None else None, ca, af,
                                    so.h if so is not None else None, co, of, nrm.h))


def _axpby_pair(out1, s1a, c1a, f1a, x1, s1b, c1b, f1b, y1, out2, s2a, c2a, f2a, x2, s2b, c2b, f2b, y2):
    """Two axpby updates in one pass (jets_axpby_pair_dev); a None scalar means "use the constant"."""
    h = lambda s: s.h if s is not None else None
    check(lib.jets_axpby_pair_dev(out1._h, h(s1a), c1a, f1a, x1._h, h(s1b), c1b, f1b, y1._h,
                                  out2._h, h(s2a), c2a, f2a, x2._h, h(s2b), c2b, f2b, y2._h))


def _sprog(steps):
    """[(out, op, a, b|None), ...] scalar operations in ONE launch (jets_scalar_prog)."""
    n = len(steps)
    arr = C.c_void_p * n
    outs = arr(*[st[0].h.value for st in steps])
    a = arr(*[st[2].h.value for st in steps])
    b = arr(*[(st[3].h.value if st[3] is not None else None) for st in steps])
    ops = "".join(st[1] for st in steps).encode()
    check(lib.jets_scalar_prog(n, outs, ops, a, b))


class LsqrGraphFused(LsqrGraph):
    """The same Golub-Kahan LSQR with the vector updates folded into the applies: u and v are kept
    UNNORMALISED (u~ = beta*u, v~ = alpha*v) so that

        u~ <- (1/alpha) A v~  - (alpha/beta) u~        one fused apply (reads v~, state, u~; writes u~)
        v~ <- (1/beta') A'u~  - (beta'/alpha) v~       one fused apply

    replace apply + axpy + scale (9 -> 4 vector passes per half iteration), ``beta = ||u~||`` and ``alpha = ||v~||``
    come out of the same launches (``jets_apply_axpby_norm``: no norm pass), the scalar recurrences run as two
    scalar programs, and the normalisations are folded into the coefficients of the x / w updates.  Same
    iterates as ``LsqrGraph`` up to rounding."""

    def __init__(self, A, b):
        self.x = x = J.zeros(J.domain(A))
        At = J.adjoint(A)
        u = b.copy()
        (beta, alpha, rho, rhobar, phibar, phi, theta, c, s, t, t1, t2, tab, tba) = (_S() for _ in range(14))
        self.alpha, self.beta = alpha, beta
        check(lib.jets_norm_dev(u._h, 2.0, beta.h))
        v = J.zeros(J.domain(A))
        _apply_axpby(v, At, u, beta, 0.0, L.COEF_INV, None, 0.0, 0)          # v~ = (1/beta) A' u~
        check(lib.jets_norm_dev(v._h, 2.0, alpha.h))
        w = J.zeros(J.domain(A))
        _axpby(w, alpha, 0.0, L.COEF_INV, v)                                   # w = v~ / alpha
        _sprog([(phibar, "+", beta, None), (rhobar, "+", alpha, None), (tab, "/", alpha, beta)])
        self._keep = (A, At, u, v, w, rho, rhobar, phibar, phi, theta, c, s, t, t1, t2, tab, tba)

        def body():
            # u~ = A v~/alpha - (alpha/beta) u~ ; beta = ||u~|| comes out of the same launch (store-epilogue partials +
            # a one-block finish): the coefficients are read when the launch starts, the norm is written after it
            _apply_axpby_norm(u, A, v, alpha, 0.0, L.COEF_INV, tab, 0.0, L.COEF_NEG, beta)
            _sprog([(tba, "/", beta, alpha)])
            _apply_axpby_norm(v, At, u, beta, 0.0, L.COEF_INV, tba, 0.0, L.COEF_NEG, alpha)   # v~ = A'u~/beta - (beta/alpha) v~ ; alpha = ||v~||
            _sprog([(rho, "h", rhobar, beta), (c, "/", rhobar, rho), (s, "/", beta, rho), (theta, "*", s, alpha),
                    (t, "*", c, alpha), (rhobar, "n", t, None), (phi, "*", c, phibar), (phibar, "*", s, phibar),
                    (t1, "/", phi, rho), (t2, "/", theta, rho), (tab, "/", alpha, beta)])
            # x += (phi/rho) w and w = v~/alpha - (theta/rho) w in ONE pass (w read once, before it is overwritten)
            _axpby_pair(x, None, 1.0, 0, x, t1, 0.0, 0, w,
                        w, alpha, 0.0, L.COEF_INV, v, t2, 0.0, L.COEF_NEG, w)

        body()  # iteration 1; builds every plan outside the capture
        check(lib.jets_graph_begin())
        try:
            body()
        finally:
            g = C.c_void_p()
            check(lib.jets_graph_end(C.byref(g)))
        self._g = g


def lsqr_graph(A, b, iters=10):
    """``iters`` iterations of LsqrGraph; returns (x, (alpha, beta))."""
    G = LsqrGraph(A, b)
    G.run(iters - 1)
    return G.result()
