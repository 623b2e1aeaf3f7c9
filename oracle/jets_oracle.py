"""CPU oracle: a numpy restatement of the Jets.jl operator-application path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The product
(``jets.jl_b200``) never imports or falls back to this module.

Parity status: **unpinned at the bit level for reductions and dense products**.  The reference
(``/root/reference/src/Jets.jl``, v1.4.1) is pure Julia; Julia is not installed in this image,
so the reference cannot be executed here and it ships no golden vectors (no seeds, no stored
outputs).  What pins this oracle is (a) a line-by-line restatement of the reference's loop
order, temporaries and quirks, each function citing the ``src/Jets.jl`` lines it follows, and
(b) every algebraic identity the reference's own ``test/runtests.jl`` asserts for this path,
restated on seeded inputs in ``tests/test_oracle_reference_identities.py``.  Arithmetic that
lives outside ``/root/reference`` -- Julia broadcast codegen (one IEEE rounding per element op,
no FMA contraction) and stdlib LinearAlgebra -> OpenBLAS ``dot``/``nrm2``/``gemv`` (version
unpinned: ``Project.toml`` has no Manifest) -- is restated with numpy elementwise ops (exact,
same roundings) and numpy ``dot``/``norm``/``@`` (summation order differs from OpenBLAS builds
by rounding only; the reference's own tests pin these to ``isapprox`` rtol=sqrt(eps)).

Conventions
-----------
* ``mul_(d, A, m)`` is Julia's ``mul!(d, A, m)``; ``A * m`` is Julia's ``A * m``; ``A.T`` or
  ``adjoint(A)`` is ``A'``; ``compose(A2, A1)`` or ``A2 @ A1`` is ``A2 ∘ A1``.
* Jet closures have the reference's signature ``f(d, m, **state)``, ``df(d, m, mo=..., **state)``,
  ``dft(m, d, mo=..., **state)`` and MUST return the output array (quirk Q11).
* Arrays are numpy arrays in Fortran (column-major) order semantics: ``vec`` and ``reshape`` use
  ``order='F'`` so linear indices agree with Julia's.
* Block index ranges are kept 1-based inclusive ``(start, stop)`` exactly as
  ``JetBSpace.indices`` (``src/Jets.jl:742-748``).
"""
from __future__ import annotations

import math
from typing import Callable, Sequence

import numpy as np

__all__ = [
    "JetSpace", "JetBSpace", "BlockArray", "Jet", "Jop", "JopNl", "JopLn", "JopAdjoint",
    "JopBlock", "JopZeroBlock", "blockop", "mul_", "adjoint", "jacobian", "jacobian_", "point",
    "point_", "state", "state_", "domain", "range_", "shape", "size", "jet", "compose",
    "nblocks", "getblock", "getblock_", "setblock_", "isblockop", "iszero", "indices", "space",
    "dot", "norm", "extrema", "fill_", "to_array", "to_matrix", "dot_product_test",
    "linearity_test", "linearization_test", "zeros", "ones", "rand", "randn", "Array", "vec",
    "close", "perfstat", "JopDiagonal", "JopPointwise", "JopStencil", "JopDense", "JopScale", "JopRestriction",
    "bmap", "PW_FUNCS", "JetSSpace", "SymmetricArray", "symspace",
]


# --------------------------------------------------------------------------------------
# L0  vector spaces                                             src/Jets.jl:5-129, 736-807
# --------------------------------------------------------------------------------------
class JetAbstractSpace:
    pass


class JetSpace(JetAbstractSpace):
    """``JetSpace(T, n...)`` -- eltype + shape metadata (src/Jets.jl:40-68)."""

    def __init__(self, T, *n):
        if len(n) == 1 and isinstance(n[0], (tuple, list)):
            n = tuple(n[0])
        self.T = np.dtype(T)
        self.n = tuple(int(k) for k in n)

    def __eq__(self, other):
        return isinstance(other, JetSpace) and self.T == other.T and self.n == other.n

    def __hash__(self):
        return hash((self.T, self.n))

    def __repr__(self):
        return f"JetSpace({self.T}, {self.n})"

    @property
    def eltype(self):
        return self.T

    @property
    def ndims(self):
        return len(self.n)

    def size(self, i=None):
        return self.n if i is None else self.n[i - 1]

    def __len__(self):
        return int(np.prod(self.n, dtype=np.int64)) if self.n else 1

    def vec(self):  # src/Jets.jl:66
        return JetSpace(self.T, len(self))

    def similar(self, *dims):  # src/Jets.jl:67-68
        if len(dims) == 1 and isinstance(dims[0], (tuple, list)):
            dims = tuple(dims[0])
        return JetSpace(self.T, *dims)


class JetBSpace(JetAbstractSpace):
    """Block space; cumulative 1-based inclusive index ranges (src/Jets.jl:736-751)."""

    def __init__(self, spaces: Sequence[JetAbstractSpace]):
        self.spaces = list(spaces)
        self.T = np.result_type(*[s.T for s in self.spaces])  # promote_type, :740-741
        self.indices = []
        stop = 0
        for s in self.spaces:  # :743-748
            start = stop + 1
            stop = start + len(s) - 1
            self.indices.append((start, stop))

    def __eq__(self, other):  # :753
        return (isinstance(other, JetBSpace) and self.spaces == other.spaces
                and self.indices == other.indices)

    def __hash__(self):
        return hash(tuple(self.indices))

    def __repr__(self):
        return f"JetBSpace({self.spaces})"

    @property
    def eltype(self):
        return self.T

    @property
    def ndims(self):
        return 1

    def size(self, i=None):  # :755
        n = (self.indices[-1][1],)
        return n if i is None else n[i - 1]

    def __len__(self):
        return self.indices[-1][1]

    def vec(self):  # :758
        return self

    def similar(self, *dims):  # :759-760
        if len(dims) == 1 and isinstance(dims[0], (tuple, list)):
            dims = tuple(dims[0])
        return JetSpace(self.T, *dims)


def indices(R, iblock):  # :780, :858
    return R.indices[iblock - 1]


def nblocks(x, i=None):  # :806-807, :860, :1074-1077
    if isinstance(x, JetBSpace):
        return len(x.spaces)
    if isinstance(x, JetAbstractSpace):
        return 1
    if isinstance(x, BlockArray):
        return len(x.indices)
    if isinstance(x, (Jet, Jop)):
        nb = (nblocks(range_(x)), nblocks(domain(x)))
        return nb if i is None else nb[i - 1]
    raise TypeError(x)


# --------------------------------------------------------------------------------------
# L1  BlockArray                                                    src/Jets.jl:809-924
# --------------------------------------------------------------------------------------
class BlockArray:
    """Vector of independently stored block arrays + their index ranges (:809-812)."""

    __array_priority__ = 100
    __array_ufunc__ = None  # numpy scalars/arrays defer to the reflected operators below

    def __init__(self, arrays, idx):
        self.arrays = list(arrays)
        self.indices = list(idx)

    @property
    def dtype(self):
        return self.arrays[0].dtype

    def __len__(self):  # :818
        return self.indices[-1][1]

    @property
    def shape(self):
        return (len(self),)

    def _find(self, i):  # O(nblocks) findfirst, :820-827   (i is 1-based)
        for j, (a, b) in enumerate(self.indices):
            if a <= i <= b:
                return j, i - a
        raise IndexError(i)

    def __getitem__(self, i):
        j, k = self._find(i)
        if isinstance(self.arrays[j], BlockArray):
            return self.arrays[j][k + 1]
        return self.arrays[j].reshape(-1, order="F")[k]

    def __setitem__(self, i, v):
        if i is Ellipsis:  # ``x .= v``
            self.assign(v) if isinstance(v, (BlockArray, np.ndarray)) else fill_(self, v)
            return
        j, k = self._find(i)
        if isinstance(self.arrays[j], BlockArray):
            self.arrays[j][k + 1] = v
            return
        flat = self.arrays[j].reshape(-1, order="F")
        flat[k] = v
        self.arrays[j][...] = flat.reshape(self.arrays[j].shape, order="F")

    def similar(self, T=None):  # :829-832
        T = self.dtype if T is None else T
        return BlockArray([a.similar(T) if isinstance(a, BlockArray) else np.empty(a.shape, dtype=T) for a in self.arrays],
                          self.indices)

    def copy(self):
        return BlockArray([a.copy() for a in self.arrays], self.indices)

    # Broadcasting (:889-911): per-block application of the fused elementwise expression.
    def _bin(self, other, f):
        return bmap(f, self, other)

    def __add__(self, o):
        return self._bin(o, lambda a, b: a + b)

    def __radd__(self, o):
        return bmap(lambda a, b: b + a, self, o)

    def __sub__(self, o):
        return self._bin(o, lambda a, b: a - b)

    def __rsub__(self, o):
        return bmap(lambda a, b: b - a, self, o)

    def __mul__(self, o):
        return self._bin(o, lambda a, b: a * b)

    def __rmul__(self, o):
        return bmap(lambda a, b: b * a, self, o)

    def __truediv__(self, o):
        return self._bin(o, lambda a, b: a / b)

    def __neg__(self):
        return bmap(lambda a: -a, self)

    def __abs__(self):
        return bmap(np.abs, self)

    def assign(self, src):
        """``dest .= src`` (copyto!, :905-911)."""
        for i in range(len(self.arrays)):
            self.arrays[i][...] = _blk(src, i, self.indices[i], self.arrays[i].shape)
        return self


def _blk(arg, iblock, rng, shp):
    """getblock(arg, S, iblock, indices): BlockArray -> its block; plain array -> A[indices]
    (allocating slice, :902); scalar -> itself (:900-903)."""
    if isinstance(arg, BlockArray):
        return arg.arrays[iblock]
    if isinstance(arg, np.ndarray) and arg.ndim >= 1:
        flat = arg.reshape(-1, order="F")[rng[0] - 1:rng[1]]
        return flat.reshape(shp, order="F")
    return arg


def bmap(f: Callable, *args):
    """Per-block map: the oracle's spelling of a Julia dot-broadcast over BlockArrays."""
    ref = next(a for a in args if isinstance(a, BlockArray))  # find_blockarray, :894-898
    out = []
    for i in range(len(ref.arrays)):
        blk = [_blk(a, i, ref.indices[i], ref.arrays[i].shape) for a in args]
        res = f(*blk)   # blocks that are BlockArrays themselves (a JetBSpace of JetBSpaces) recurse through their operators
        out.append(res if isinstance(res, BlockArray) else np.asarray(res))
    return BlockArray(out, ref.indices)


# --------------------------------------------------------------------------------------
# Symmetric spaces / arrays                                         src/Jets.jl:405-516
# --------------------------------------------------------------------------------------
class JetSSpace(JetAbstractSpace):
    """``JetSSpace(T, n, M, map)`` (:408-441): logical size n, stored parent of size M; ``map``
    takes the (1-based) index tuple of a position beyond the parent and returns the index of the
    stored element whose conjugate it is."""

    def __init__(self, T, n, M, map):
        self.T = np.dtype(T)
        self.n = tuple(int(k) for k in n)
        self.M = tuple(int(k) for k in M)
        self.map = map

    def __eq__(self, o):
        return isinstance(o, JetSSpace) and self.T == o.T and self.n == o.n and self.M == o.M and self.map is o.map

    def __hash__(self):
        return hash((self.T, self.n, self.M))

    eltype = property(lambda s: s.T)
    ndims = property(lambda s: len(s.n))

    def size(self, i=None):  # :437
        return self.n if i is None else self.n[i - 1]

    def __len__(self):
        return int(np.prod(self.n, dtype=np.int64))

    def similar(self, *dims):  # :441-442
        if len(dims) == 1 and isinstance(dims[0], (tuple, list)):
            dims = tuple(dims[0])
        return JetSSpace(self.T, dims, self.M, self.map)


def symspace():  # :443
    return None


class SymmetricArray:
    """``SymmetricArray`` (:445-484): parent ``A`` (size M), logical size ``n``; 1-based indexing;
    ``x[I]`` beyond the parent in any dimension is ``conj(A[map(I)])`` (:455-462) and assignment
    there stores ``conj(v)`` (:470-478).  Linear indices are column-major over ``n`` (:465-468)."""

    def __init__(self, A, n, map):
        self.A = A
        self.n = tuple(n)
        self.map = map

    dtype = property(lambda s: s.A.dtype)
    shape = property(lambda s: s.n)

    def parent(self):  # :449
        return self.A

    def _cart(self, I):
        if isinstance(I, tuple):
            return tuple(int(k) for k in I)
        return tuple(int(k) + 1 for k in np.unravel_index(int(I) - 1, self.n, order="F"))

    def __getitem__(self, I):
        I = self._cart(I)
        for d in range(len(self.n)):
            if I[d] > self.A.shape[d]:
                J = self.map(I)
                return np.conj(self.A[tuple(j - 1 for j in J)])
        return self.A[tuple(i - 1 for i in I)]

    def __setitem__(self, I, v):
        I = self._cart(I)
        for d in range(len(self.n)):
            if I[d] > self.A.shape[d]:
                J = self.map(I)
                self.A[tuple(j - 1 for j in J)] = np.conj(v)
                return
        self.A[tuple(i - 1 for i in I)] = v

    def full(self):
        """collect(x): every logical entry through getindex."""
        out = np.empty(self.n, dtype=self.dtype, order="F")
        for I in np.ndindex(*self.n):
            out[I] = self[tuple(i + 1 for i in I)]
        return out

    # broadcast acts on the parents (:486-508)
    def _b(self, o, f):
        return SymmetricArray(f(self.A, o.A if isinstance(o, SymmetricArray) else o), self.n, self.map)

    def __add__(self, o): return self._b(o, lambda a, b: a + b)
    __radd__ = __add__
    def __sub__(self, o): return self._b(o, lambda a, b: a - b)
    def __mul__(self, o): return self._b(o, lambda a, b: a * b)
    __rmul__ = __mul__

    def assign(self, o):  # copyto!(dest::SymmetricArray, bc) :503-507
        self.A[...] = o.A if isinstance(o, SymmetricArray) else o
        return self

    def similar(self):  # :480-482
        return SymmetricArray(np.empty_like(self.A), self.n, self.map)


def space(x, iblock=None):
    """space(x::AbstractArray) :126; space(x::BlockArray) :814; space(R, iblock) :799."""
    if isinstance(x, SymmetricArray):  # :447
        return JetSSpace(x.dtype, x.n, x.A.shape, x.map)
    if isinstance(x, JetBSpace):
        return x.spaces[iblock - 1]
    if isinstance(x, BlockArray):
        return JetBSpace([space(a) for a in x.arrays])
    return JetSpace(x.dtype, *x.shape)


def _factory(kind):
    def make(R, rng=None):
        if isinstance(R, JetBSpace):  # :922-924
            return BlockArray([make(s, rng) for s in R.spaces], R.indices)
        if isinstance(R, JetSSpace):  # :510-513: the factory acts on the parent shape M
            return SymmetricArray(make(JetSpace(R.T, *R.M), rng), R.n, R.map)
        shp = R.n
        if kind == "zeros":
            return np.zeros(shp, dtype=R.T, order="F")
        if kind == "ones":
            return np.ones(shp, dtype=R.T, order="F")
        if kind == "Array":
            return np.empty(shp, dtype=R.T, order="F")
        g = rng if rng is not None else np.random.default_rng()
        draw = g.random if kind == "rand" else g.standard_normal
        if np.issubdtype(R.T, np.complexfloating):
            real_t = np.float32 if R.T == np.complex64 else np.float64
            x = draw(shp).astype(real_t) + 1j * draw(shp).astype(real_t)
            return np.asfortranarray(x.astype(R.T))
        return np.asfortranarray(draw(shp).astype(R.T))
    make.__name__ = kind
    return make


zeros = _factory("zeros")
ones = _factory("ones")
rand = _factory("rand")
randn = _factory("randn")
Array = _factory("Array")


def getblock(x, i, j=None):
    if j is not None:
        return _getblock_op(x, i, j)
    if isinstance(x, BlockArray):  # :914 -- a reference, not a copy
        return x.arrays[i - 1]
    return x  # :918


def getblock_(x, iblock, xblock):  # :915, :919
    xblock[...] = x.arrays[iblock - 1] if isinstance(x, BlockArray) else x
    return xblock


def setblock_(x, iblock, xblock):  # :916, :920
    if isinstance(x, BlockArray):
        x.arrays[iblock - 1][...] = xblock
        return x.arrays[iblock - 1]
    x[...] = xblock
    return x


def norm(x, p=2):
    """norm(x::BlockArray, p) (:834-848): per-block stdlib norm, combined per p.  A SymmetricArray
    has no norm method of its own: the generic AbstractArray norm iterates the logical array."""
    if isinstance(x, SymmetricArray):
        return _norm1(x.full(), p)
    if not isinstance(x, BlockArray):
        return _norm1(x, p)
    if p == math.inf:
        return max(_norm1(a, p) for a in x.arrays)
    if p == -math.inf:
        return min(_norm1(a, p) for a in x.arrays)
    if p == 1 or p == 0:
        s = 0.0
        for a in x.arrays:
            s = s + _norm1(a, p)
        return s
    rt = np.float32 if x.dtype in (np.float32, np.complex64) else np.float64
    _p = rt(p)
    s = None
    for a in x.arrays:  # mapreduce(_x->norm(_x,p)^_p, +, arrays)
        t = rt(_norm1(a, p)) ** _p
        s = t if s is None else s + t
    return s ** (rt(1) / _p)


def _norm1(a, p):
    if isinstance(a, BlockArray):   # norm(_x, p) of a block that is a BlockArray itself (:834-848, recursively)
        return norm(a, p)
    v = np.asarray(a).reshape(-1)
    if p == 2:
        return np.linalg.norm(v)
    if p == math.inf:
        return np.max(np.abs(v)) if v.size else 0.0
    if p == -math.inf:
        return np.min(np.abs(v)) if v.size else 0.0
    if p == 1:
        return np.sum(np.abs(v))
    if p == 0:
        return float(np.count_nonzero(v))
    return np.sum(np.abs(v) ** p) ** (1.0 / p)


def dot(x, y):
    """dot(x::BlockArray, y) (:850-856): a = zero(T); a += dot(x_i, y_i) in block order;
    conjugates the first argument."""
    if isinstance(x, BlockArray):
        a = x.dtype.type(0)
        for xi, yi in zip(x.arrays, y.arrays):
            a = a + (dot(xi, yi) if isinstance(xi, BlockArray) else np.vdot(xi.reshape(-1, order="F"), yi.reshape(-1, order="F")))
        return a
    return np.vdot(np.asarray(x).reshape(-1, order="F"), np.asarray(y).reshape(-1, order="F"))


def extrema(x):  # :870-878
    if not isinstance(x, BlockArray):
        return np.min(x), np.max(x)
    mn, mx = extrema(x.arrays[0])
    for a in x.arrays[1:]:
        _mn, _mx = extrema(a)
        if _mn < mn:
            mn = _mn
        if _mx > mx:
            mx = _mx
    return mn, mx


def fill_(x, a):  # :880-885
    if isinstance(x, BlockArray):
        for b in x.arrays:
            b[...] = a
    else:
        x[...] = a
    return x


def to_array(x):
    """convert(Array, x::BlockArray) (:862-868)."""
    if not isinstance(x, BlockArray):
        return np.asarray(x)
    out = np.empty(len(x), dtype=x.dtype)
    for (a, b), blk in zip(x.indices, x.arrays):
        out[a - 1:b] = to_array(blk).reshape(-1, order="F")   # vec(x.arrays[i]); a nested block flattens the same way
    return out


def reshape(x, R):
    """reshape(x, R): plain space :38; block space -> per-block *views* of flat x (:1112);
    BlockArray -> itself after a length check (:1115-1118)."""
    if isinstance(R, JetBSpace):
        if isinstance(x, BlockArray):
            if len(x) != len(R):
                raise ValueError("dimension mismatch, unable to reshape block array")
            return x
        flat = x.reshape(-1, order="F") if x.flags.f_contiguous else None
        if flat is None or not np.shares_memory(flat, x):
            flat = x.reshape(-1)  # 1-D input
        blocks = []
        for (a, b), s in zip(R.indices, R.spaces):
            v = flat[a - 1:b]
            blocks.append(reshape(v, s))   # reshape(view(x, indices), R.spaces[i]): recursive for a nested block space
        return BlockArray(blocks, R.indices)
    if isinstance(x, BlockArray):
        return to_array(x).reshape(R.n, order="F")
    return x.reshape(R.n, order="F")


def _copy(x):
    return x.copy() if isinstance(x, BlockArray) else np.array(x, copy=True, order="F")


def _assign(dst, src):
    """``dst .= src``."""
    if isinstance(dst, BlockArray):
        return dst.assign(src)
    dst[...] = to_array(src).reshape(dst.shape, order="F") if isinstance(src, BlockArray) else src
    return dst


# --------------------------------------------------------------------------------------
# L2  Jet                                                          src/Jets.jl:131-192
# --------------------------------------------------------------------------------------
def jet_missing(*a, **k):  # :131
    raise NotImplementedError("not implemented")


class Jet:
    """Plugin record (:133-142) with the constructor's defaulting rules (:170-188)."""

    def __init__(self, *, dom, rng, f=jet_missing, df=jet_missing, dft=jet_missing,
                 upstate=None, s=None):
        if f is jet_missing and df is jet_missing:  # :178-180
            raise ValueError("must set at-least one of f! and df!")
        if f is jet_missing:  # :181-183
            f = df
        if dft is jet_missing:  # :184-186
            dft = df
        self.dom, self.rng = dom, rng
        self.f, self.df, self.dft = f, df, dft
        self.upstate = upstate if upstate is not None else (lambda m, s: None)
        self.mo = np.empty((0,) * dom.ndims, dtype=dom.T)  # :187
        self.s = dict(s or {})

    def copy(self, copymo=True):  # :230 -- deepcopy(jet.s)
        import copy as _c
        j = Jet.__new__(Jet)
        j.dom, j.rng, j.f, j.df, j.dft, j.upstate = (self.dom, self.rng, self.f, self.df,
                                                      self.dft, self.upstate)
        j.mo = _copy(self.mo) if copymo else self.mo
        j.s = _c.deepcopy(self.s)
        return j


# --------------------------------------------------------------------------------------
# L3  operators                                                    src/Jets.jl:194-403
# --------------------------------------------------------------------------------------
class Jop:
    def __mul__(self, m):  # :399   A*m = mul!(zeros(range(A)), A, m)
        if isinstance(m, (np.ndarray, BlockArray)):
            return mul_(zeros(range_(self)), self, m)
        return NotImplemented

    def __rmul__(self, a):  # :1161-1164   a*A
        if np.isscalar(a):
            return scalar_mul(a, self)
        return NotImplemented

    def __matmul__(self, other):
        return compose(self, other)

    def __rmatmul__(self, other):
        return compose(other, self)

    def __add__(self, other):
        return op_sum(self, other, +1)

    def __radd__(self, other):
        return op_sum(other, self, +1)

    def __sub__(self, other):
        return op_sum(self, other, -1)

    def __rsub__(self, other):
        return op_sum(other, self, -1)

    @property
    def T(self):
        return adjoint(self)

    __array_priority__ = 200
    __array_ufunc__ = None


class JopNl(Jop):
    def __init__(self, jet_=None, **kw):
        self.jet = jet_ if jet_ is not None else Jet(**kw)  # :207

    def copy(self, copymo=True):
        return JopNl(self.jet.copy(copymo))


class JopLn(Jop):
    def __new__(cls, a=None, mo=None, **kw):
        if isinstance(a, (JopLn, JopAdjoint)) and mo is None:  # :223, :235
            return a
        return super().__new__(cls)

    def __init__(self, a=None, mo=None, **kw):
        if isinstance(a, (JopLn, JopAdjoint)):
            return
        if isinstance(a, JopNl):  # :224
            a = a.jet
        if a is None:
            a = Jet(**kw)  # :221
        if mo is not None:  # :212
            point_(a, mo)
        self.jet = a

    def copy(self, copymo=True):
        return JopLn(self.jet.copy(copymo))


class JopAdjoint(Jop):
    def __init__(self, op):
        self.op = op

    @property
    def jet(self):  # :309
        return self.op.jet

    def copy(self, copymo=True):
        return JopAdjoint(self.op.copy(copymo))


def jet(A):
    return A if isinstance(A, Jet) else A.jet


def domain(A):  # :242, :319, :322, :325
    if isinstance(A, np.ndarray):
        return JetSpace(A.dtype, A.shape[1])
    if isinstance(A, JopAdjoint):
        return range_(A.op)
    return jet(A).dom


def range_(A):  # :249, :320, :323, :326
    if isinstance(A, np.ndarray):
        return JetSpace(A.dtype, A.shape[0])
    if isinstance(A, JopAdjoint):
        return domain(A.op)
    return jet(A).rng


def eltype(A):  # :256
    return np.result_type(domain(A).T, range_(A).T)


def shape(A, i=None):  # :328-345
    if isinstance(A, np.ndarray):
        s = ((A.shape[0],), (A.shape[1],))
    else:
        s = (range_(A).size(), domain(A).size())
    return s if i is None else s[0 if i == 1 else 1]


def size(A, i=None):  # :354-355
    s = (len(range_(A)), len(domain(A)))
    return s if i is None else s[0 if i == 1 else 1]


def state(A, key=None):  # :264-265, :313-314, composite lookup :607-623
    j = jet(A)
    if key is None:
        return j.s
    if j.f is JetComposite_f and key not in j.s:
        hits = [op for op in j.s["ops"] if key in state(op)]
        if not hits:
            raise KeyError(f"key {key} does not exist in the state of the composite operator")
        if len(hits) > 1:
            raise KeyError(f"ambiguous: key {key} exists in more than one operator in the composition")
        return state(hits[0], key)
    return j.s[key]


def state_(A, s):  # :272
    jet(A).s.update(s)
    return A


def point(A):  # :288, :311-312
    return jet(A).mo


def point_(j, mo):
    """point!(jet, mo): leaf :297-301; composite :578-589; sum :710-715; block :1059-1066."""
    j = jet(j)
    if j.f is JetComposite_f:
        j.mo = mo
        ops = j.s["ops"]
        _m = _copy(mo)
        for i in range(len(ops) - 1, -1, -1):
            point_(jet(ops[i]), _m)
            if i > 0:
                _m = mul_(zeros(range_(ops[i])), ops[i], _m)
        return j
    if j.f is JetSum_f:
        for op in j.s["ops"]:
            point_(jet(op), mo)
        return j
    if j.f is JetBlock_f:
        ops = j.s["ops"]
        for icol in range(ops.shape[1]):
            for irow in range(ops.shape[0]):
                point_(jet(ops[irow, icol]), getblock(mo, icol + 1))
        return j
    if j.f is JetVec_f:
        point_(jet(j.s["op"]), reshape(mo, domain(j.s["op"])))
        j.mo = mo
        return j
    j.mo = mo  # by reference, :298
    j.upstate(mo, j.s)
    return j


def jacobian_(F, mo):
    """jacobian!(F, mo): shares (and mutates) the underlying jet (:364-366)."""
    if isinstance(F, (JopLn, JopAdjoint, np.ndarray)):
        return F
    return JopLn(jet(F), mo)


def jacobian(F, mo):
    """jacobian(F, mo) = jacobian!(copy(F,false), copy(mo)) (:374-375)."""
    if isinstance(F, np.ndarray):
        return F.copy()
    return jacobian_(F.copy(False), _copy(mo))


def adjoint(A):  # :382-383
    if isinstance(A, JopAdjoint):
        return A.op
    if isinstance(A, JopLn):
        return JopAdjoint(A)
    if isinstance(A, np.ndarray):
        return A.conj().T
    raise TypeError("adjoint is defined for JopLn/JopAdjoint only")


def mul_(d, A, m):
    """mul! dispatch table (:390-392).  State is splatted as keyword arguments."""
    if isinstance(A, JopNl):
        return A.jet.f(d, m, **A.jet.s)
    if isinstance(A, JopLn):
        return A.jet.df(d, m, mo=A.jet.mo, **A.jet.s)
    if isinstance(A, JopAdjoint):
        if not isinstance(A.op, JopLn):
            raise TypeError("mul! is undefined for the adjoint of a nonlinear operator")
        return A.jet.dft(d, m, mo=A.jet.mo, **A.jet.s)
    if isinstance(A, np.ndarray):
        d[...] = A @ m
        return d
    raise TypeError(A)


def close(A):  # :290, :317, :591-595, :717-721, :1120-1124
    j = jet(A)
    if j.f in (JetComposite_f, JetSum_f):
        for op in j.s["ops"]:
            close(op)
        return None
    if j.f is JetBlock_f:
        for op in j.s["ops"].reshape(-1):
            close(op)
        return None
    c = j.s.get("_close")
    return c(j) if c is not None else False


def perfstat(A):  # :281, :316, :597-605, :723-731
    j = jet(A)
    if j.f in (JetComposite_f, JetSum_f):
        s = None
        for op in j.s["ops"]:
            s = perfstat(op)
            if s is not None:
                break
        return s
    p = j.s.get("_perfstat")
    return p(j) if p is not None else None


# --------------------------------------------------------------------------------------
# L4  composition  f ∘ g                                           src/Jets.jl:518-623
# --------------------------------------------------------------------------------------
def JetComposite_f(d, m, *, ops, **kw):  # :524-528
    x = m
    for i in range(len(ops) - 1, -1, -1):  # ops[n] first ... ops[1] last
        x = mul_(zeros(range_(ops[i])), ops[i], x)
    return _assign(d, x)


def JetComposite_df(d, m, *, ops, **kw):  # :530-534
    x = m
    for i in range(len(ops) - 1, -1, -1):
        L = JopLn(ops[i])
        x = mul_(zeros(range_(L)), L, x)
    return _assign(d, x)


def JetComposite_dft(m, d, *, ops, **kw):  # :536-540   ops[1]' first ... ops[n]' last
    x = d
    for i in range(len(ops)):
        L = JopLn(ops[i])
        x = mul_(zeros(domain(L)), adjoint(L), x)
    return _assign(m, x)


def JetComposite(ops):  # :522
    ops = tuple(ops)
    return Jet(f=JetComposite_f, df=JetComposite_df, dft=JetComposite_dft,
               dom=domain(ops[-1]), rng=range_(ops[0]), s={"ops": ops})


def _jops_comp(op):  # :542-550
    if isinstance(op, (JopLn, JopNl)) and op.jet.f is JetComposite_f:
        return tuple(op.jet.s["ops"])
    if isinstance(op, JopAdjoint) and isinstance(op.op, JopLn) and op.jet.f is JetComposite_f:
        return tuple(JopAdjoint(o) if not isinstance(o, JopAdjoint) else o.op
                     for o in reversed(op.op.jet.s["ops"]))
    return (op,)


def _wrap_matrix(A):  # :573-576
    def _df(d, m, *, A, **kw):
        d[...] = A @ m
        return d

    def _dft(m, d, *, A, **kw):
        m[...] = A.conj().T @ d
        return m
    return JopLn(dom=domain(A), rng=range_(A), df=_df, dft=_dft, s={"A": A})


def compose(A2, A1):
    """A2 ∘ A1 (:569-576): flattens nested composites; all-linear -> JopLn else JopNl."""
    if isinstance(A2, np.ndarray) and isinstance(A1, np.ndarray):
        return A2 @ A1
    if isinstance(A1, np.ndarray):
        A1 = _wrap_matrix(A1)
    if isinstance(A2, np.ndarray):
        A2 = _wrap_matrix(A2)
    ops = _jops_comp(A2) + _jops_comp(A1)
    lin = isinstance(A2, (JopLn, JopAdjoint)) and isinstance(A1, (JopLn, JopAdjoint))
    return JopLn(JetComposite(ops)) if lin else JopNl(JetComposite(ops))


# --------------------------------------------------------------------------------------
# L4  sums  f ± g                                                  src/Jets.jl:625-731
# --------------------------------------------------------------------------------------
def _sgn_apply(sgn, d, t):
    """broadcast!(sgn, d, d, t): d .= d ± t."""
    if isinstance(d, BlockArray):
        for a, b in zip(d.arrays, t.arrays):
            a[...] = a + b if sgn > 0 else a - b
    else:
        d[...] = d + t if sgn > 0 else d - t
    return d


def JetSum_f(d, m, *, ops, sgns, **kw):  # :630-637
    fill_(d, 0)
    _d = zeros(range_(ops[0]))
    for op, s in zip(ops, sgns):
        _sgn_apply(s, d, mul_(_d, op, m))
    return d


def JetSum_df(d, m, *, ops, sgns, **kw):  # :639-646
    fill_(d, 0)
    _d = zeros(range_(ops[0]))
    for op, s in zip(ops, sgns):
        _sgn_apply(s, d, mul_(_d, JopLn(op), m))
    return d


def JetSum_dft(m, d, *, ops, sgns, **kw):  # :648-655
    fill_(m, 0)
    _m = zeros(domain(ops[0]))
    for op, s in zip(ops, sgns):
        _sgn_apply(s, m, mul_(_m, adjoint(JopLn(op)), d))
    return m


def JetSum(ops, sgns):  # :628
    return Jet(f=JetSum_f, df=JetSum_df, dft=JetSum_dft, dom=domain(ops[0]),
               rng=range_(ops[0]), s={"ops": tuple(ops), "sgns": tuple(sgns)})


def _is_sum(op):
    return op.jet.f is JetSum_f and (isinstance(op, (JopLn, JopNl)) or
                                     (isinstance(op, JopAdjoint) and isinstance(op.op, JopLn)))


def _jops_sum(op):  # :657-665
    if _is_sum(op):
        ops = op.jet.s["ops"]
        if isinstance(op, JopAdjoint):
            return tuple(JopAdjoint(o) if not isinstance(o, JopAdjoint) else o.op for o in ops)
        return tuple(ops)
    return (op,)


def _sgns(op, r):  # flipsgn :667-671, sgns :673-676
    if _is_sum(op):
        return tuple(s * r for s in op.jet.s["sgns"])
    return (r,)


def op_sum(A2, A1, sign):
    """A2 + A1 / A2 - A1 (:689-708)."""
    if isinstance(A1, np.ndarray):
        A1 = _wrap_matrix(A1)
    if isinstance(A2, np.ndarray):
        A2 = _wrap_matrix(A2)
    ops = _jops_sum(A2) + _jops_sum(A1)
    sg = _sgns(A2, +1) + _sgns(A1, sign)
    lin = isinstance(A2, (JopLn, JopAdjoint)) and isinstance(A1, (JopLn, JopAdjoint))
    return JopLn(JetSum(ops, sg)) if lin else JopNl(JetSum(ops, sg))


# --------------------------------------------------------------------------------------
# L4  block operator                                               src/Jets.jl:926-1124
# --------------------------------------------------------------------------------------
def _iadd(dst, src):
    dst[...] = dst + src


def JetBlock_f(d, m, *, ops, dom, rng, **kw):  # :988-1008
    nr, nc = ops.shape
    dtmp = zeros(range_(ops[0, 0])) if nc > 1 else None
    for ir in range(nr):
        _d = getblock(d, ir + 1)
        if nc > 1 and dtmp.shape != range_(ops[ir, 0]).size():
            dtmp = zeros(range_(ops[ir, 0]))
        for ic in range(nc):
            _m = getblock(m, ic + 1)
            if nc > 1:
                _iadd(_d, mul_(dtmp, ops[ir, ic], _m))  # NOTE: _d is never zeroed (Q1)
            else:
                mul_(_d, ops[ir, ic], m)  # whole m, not _m (Q3)
    return d


def JetBlock_df(d, m, *, ops, dom, rng, **kw):  # :1010-1032
    nr, nc = ops.shape
    dtmp = zeros(range_(ops[0, 0])) if nc > 1 else None
    for ir in range(nr):
        _d = getblock(d, ir + 1)
        if nc > 1 and dtmp.shape != range_(ops[ir, 0]).size():
            dtmp = zeros(range_(ops[ir, 0]))
        for ic in range(nc):
            _m = getblock(m, ic + 1)
            if not iszero(ops[ir, ic]):  # zero blocks are skipped (Q2)
                if nc > 1:
                    _iadd(_d, mul_(dtmp, JopLn(ops[ir, ic]), _m))
                else:
                    mul_(_d, JopLn(ops[ir, ic]), _m)
    return d


def JetBlock_dft(m, d, *, ops, dom, rng, **kw):  # :1034-1057
    nr, nc = ops.shape
    mtmp = zeros(domain(ops[0, 0])) if nr > 1 else None
    for ic in range(nc):
        _m = getblock(m, ic + 1)
        if nr > 1:
            _m[...] = 0  # the adjoint DOES zero its output
            if mtmp.shape != domain(ops[0, ic]).size():
                mtmp = zeros(domain(ops[0, ic]))
        for ir in range(nr):
            _d = getblock(d, ir + 1)
            if not iszero(ops[ir, ic]):
                if nr > 1:
                    _iadd(_m, mul_(mtmp, adjoint(JopLn(ops[ir, ic])), _d))
                else:
                    mul_(_m, adjoint(JopLn(ops[ir, ic])), _d)
    return m


def JetBlock(ops, dadom=False, **kw):  # :926-930
    nr, nc = ops.shape
    dom = (domain(ops[0, 0]) if (nc == 1 and not dadom)
           else JetBSpace([domain(ops[0, i]) for i in range(nc)]))
    rng = JetBSpace([range_(ops[i, 0]) for i in range(nr)])
    s = {"ops": ops, "dom": dom, "rng": rng}
    s.update(kw)
    return Jet(f=JetBlock_f, df=JetBlock_df, dft=JetBlock_dft, dom=dom, rng=rng, s=s)


def JopBlock(ops, **kw):
    """JopBlock (:931-933): a matrix (list of rows / 2-D object array) or a vector of ops
    (a vector is a single block *column*, :933)."""
    arr = _as_op_matrix(ops)
    lin = all(isinstance(o, (JopLn, JopAdjoint)) for o in arr.reshape(-1))
    j = JetBlock(arr, **kw)
    return JopLn(j) if lin else JopNl(j)


blockop = JopBlock  # the @blockop macro, :953-986


def _as_op_matrix(ops):
    if isinstance(ops, np.ndarray) and ops.dtype == object:
        return ops.reshape(-1, 1) if ops.ndim == 1 else ops
    if len(ops) and isinstance(ops[0], (list, tuple)):
        nr, nc = len(ops), len(ops[0])
        a = np.empty((nr, nc), dtype=object)
        for i in range(nr):
            for k in range(nc):
                a[i, k] = ops[i][k]
        return a
    a = np.empty((len(ops), 1), dtype=object)
    for i, o in enumerate(ops):
        a[i, 0] = o
    return a


def JopZeroBlock_df(d, m, **kw):  # :942
    d[...] = 0
    return d


def JopZeroBlock(dom, rng):  # :941
    return JopLn(df=JopZeroBlock_df, dom=dom, rng=rng)


def iszero(A):  # :949-951
    return jet(A).df is JopZeroBlock_df and jet(A).f is JopZeroBlock_df


def isblockop(A):  # :1097-1098
    return isinstance(A, Jop) and jet(A).f is JetBlock_f


def _getblock_op(A, i, j):
    """getblock(A, i, j) (:1085-1110)."""
    if isinstance(A, JopAdjoint):  # :1088
        return adjoint(_getblock_op(A.op, j, i))
    jt = jet(A)
    if jt.f is JetBlock_f:
        blk = jt.s["ops"][i - 1, j - 1]
    elif jt.f is JetComposite_f:  # :1100-1110
        parts = [_getblock_op(op, i, j) if isblockop(op) else op for op in jt.s["ops"]]
        blk = parts[0]
        for p in parts[1:]:
            blk = compose(blk, p)
    else:
        raise TypeError("not a block operator")
    if isinstance(A, JopLn):  # :1086
        return JopLn(blk)
    return blk  # :1087


# --------------------------------------------------------------------------------------
# L4  vec(A), a*A                                         src/Jets.jl:1126-1164
# --------------------------------------------------------------------------------------
def JetVec_f(d, m, *, op, **kw):  # :1134
    mul_(reshape(d, range_(op)), op, reshape(m, domain(op)))
    return d


def JetVec_df(d, m, *, op, **kw):  # :1135
    mul_(reshape(d, range_(op)), JopLn(op), reshape(m, domain(op)))
    return d


def JetVec_dft(m, d, *, op, **kw):  # :1136
    mul_(reshape(m, domain(op)), adjoint(JopLn(op)), reshape(d, range_(op)))
    return m


def vec(x):
    """vec(A::Jop) (:1129-1154); vec(R) for spaces; vec(x) for arrays."""
    if isinstance(x, JetAbstractSpace):
        return x.vec()
    if isinstance(x, Jop):
        if domain(x).ndims == 1 and range_(x).ndims == 1:  # :1130
            return x
        j = Jet(f=JetVec_f, df=JetVec_df, dft=JetVec_dft, dom=domain(x).vec(),
                rng=range_(x).vec(), s={"op": x})
        return JopLn(j) if isinstance(x, (JopLn, JopAdjoint)) else JopNl(j)
    if isinstance(x, BlockArray):
        return x
    return x.reshape(-1, order="F")


def _constdiag_df(d, m, *, a, **kw):  # :1159
    d[...] = a * m
    return d


def _constdiag_dft(m, d, *, a, **kw):  # :1160
    m[...] = np.conj(a) * d
    return m


def scalar_mul(a, A):
    """a*A (:1161-1164).  The reference builds the scalar op on domain(A) for both spaces
    (quirk Q6: only valid for square A)."""
    _a = JopLn(dom=domain(A), rng=domain(A), df=_constdiag_df, dft=_constdiag_dft, s={"a": a})
    return compose(_a, A)


# --------------------------------------------------------------------------------------
# L5  utilities                                                  src/Jets.jl:1166-1282
# --------------------------------------------------------------------------------------
def to_matrix(A):
    """convert(Array, A::Jop) (:1174-1185): column-by-column materialisation."""
    m = zeros(domain(A))
    d = zeros(range_(A))
    nr, nc = size(A)
    B = np.zeros((nr, nc), dtype=eltype(A))
    for icol in range(nc):
        fill_(m, 0)
        fill_(d, 0)
        if isinstance(m, BlockArray):
            m[icol + 1] = 1
        else:
            flat = m.reshape(-1, order="F")
            flat[icol] = 1
            m[...] = flat.reshape(m.shape, order="F")
        B[:, icol] = vec(to_array(mul_(d, A, m)) if isinstance(d, BlockArray) else mul_(d, A, m))
    return B


def _had(a, b):
    return a * b


def dot_product_test(op, m, d, mmask=None, dmask=None):
    """(:1211-1226) lhs = <mmask.*m, A'(dmask.*d)>, rhs = <A(mmask.*m), dmask.*d>."""
    if not isinstance(op, JopLn):
        raise TypeError("dot_product_test accepts JopLn only (Q8)")
    mmask = ones(domain(op)) if mmask is None or len(mmask) == 0 else mmask
    dmask = ones(range_(op)) if dmask is None or len(dmask) == 0 else dmask
    ds = op * _had(mmask, m)
    ms = adjoint(op) * _had(dmask, d)
    lhs = dot(_had(mmask, m), ms)
    rhs = dot(ds, _had(dmask, d))
    if np.iscomplexobj(lhs) and np.iscomplexobj(rhs):
        return lhs, rhs
    return np.real(lhs), np.real(rhs)


def linearity_test(A, m1=None, m2=None, rng=None):
    """(:1276-1282); default vectors are -2*rand (quirk Q7)."""
    g = rng if rng is not None else np.random.default_rng()
    m1 = -1 * 2 * rand(domain(A), g) if m1 is None else m1
    m2 = -1 * 2 * rand(domain(A), g) if m2 is None else m2
    lhs = A * (m1 + m2)
    rhs = (A * m1) + (A * m2)
    return lhs, rhs


def linearization_test(F, mo, mu=(1.0, 0.5, 0.25, 0.125, 0.0625, 0.03125), dm=None,
                       mmask=None, dmask=None, rng=None):
    """(:1235-1266) Taylor-remainder convergence ratios."""
    g = rng if rng is not None else np.random.default_rng()
    mmask = ones(domain(F)) if mmask is None else mmask
    dmask = ones(range_(F)) if dmask is None else dmask
    dm = mmask * (-1 + 2 * rand(domain(F), g)) if dm is None else dm * mmask
    Fo = F * mo
    Jo = jacobian_(F, mo)
    Jodm = Jo * dm
    mu = np.array(sorted(mu, reverse=True), dtype=np.asarray(to_array(mo)).dtype)
    phi = np.zeros(len(mu))
    muobs = np.zeros(len(mu) - 1)
    muexp = np.zeros(len(mu) - 1)
    for i in range(len(mu)):
        d_lin = Fo + mu[i] * Jodm
        d_non = F * (mo + mu[i] * dm)
        phi[i] = norm(dmask * (d_non - d_lin))
        if i > 0:
            muobs[i - 1] = phi[i - 1] / phi[i]
            muexp[i - 1] = (mu[i - 1] / mu[i]) ** 2
    return muobs, muexp


# --------------------------------------------------------------------------------------
# Primitive registry (the build's leaf operators) expressed as oracle Jets.
# Diagonal / pointwise / dense follow the reference's fixtures (test/runtests.jl:3-33) and
# doc example (docs/src/index.md:110-113).  The stencil has NO reference definition (it lives
# in the un-vendored JetPack.jl): its semantics are defined here and pinned by
# dot_product_test + to_matrix in tests -- "parity unpinned" for that primitive.
# Each elementwise operation below is a single IEEE rounding, evaluated left to right, which
# is what Julia's fused broadcast does; the CUDA kernels are compiled with -fmad=false so the
# device result is bit-identical on these paths.
# --------------------------------------------------------------------------------------
def JopDiagonal(w):
    """d = w .* m ; adjoint m = conj(w) .* d   (fixture JopFoo, test/runtests.jl:3-8)."""
    def _df(d, m, *, diagonal, **kw):
        d[...] = diagonal * m
        return d

    def _dft(m, d, *, diagonal, **kw):
        m[...] = np.conj(diagonal) * d
        return m
    spc = JetSpace(w.dtype, *w.shape)
    return JopLn(df=_df, dft=_dft, dom=spc, rng=spc, s={"diagonal": w})


def JopScale(T, n, a):
    """d = a*m ; adjoint conj(a)*d   (_constdiag_df!, src/Jets.jl:1159-1160)."""
    spc = JetSpace(T, *((n,) if np.isscalar(n) else tuple(n)))
    a = np.dtype(T).type(a)
    return JopLn(dom=spc, rng=spc, df=_constdiag_df, dft=_constdiag_dft, s={"a": a})


# phi(x; p), phi'(x; p).  Evaluation order matters for bit parity: see kernels (fused_ops.cuh).
PW_FUNCS = {
    "square": (lambda x, p: x * x, lambda x, p: x.dtype.type(2) * x),
    "power": (lambda x, p: np.power(x, x.dtype.type(p)),
              lambda x, p: x.dtype.type(p) * np.power(x, x.dtype.type(p - 1.0))),
    "exp": (lambda x, p: np.exp(x), lambda x, p: np.exp(x)),
    "sin": (lambda x, p: np.sin(x), lambda x, p: np.cos(x)),
    "tanh": (lambda x, p: np.tanh(x), lambda x, p: x.dtype.type(1) - np.tanh(x) * np.tanh(x)),
    "log": (lambda x, p: np.log(x), lambda x, p: x.dtype.type(1) / x),
    "atan": (lambda x, p: np.arctan(x), lambda x, p: x.dtype.type(1) / (x.dtype.type(1) + x * x)),
}


def JopRestriction(T, n, indices):
    """d = m[indices]; adjoint m .= 0, m[indices] = d.  JetPack-style restriction (JetPack.jl is
    un-vendored: the definition is this build's, pinned by dot_product_test and convert(Array, A)).
    ``indices`` are 1-based and unique, as a Julia caller would pass them."""
    idx = np.asarray(indices, dtype=np.int64) - 1
    assert idx.size == np.unique(idx).size and (idx.size == 0 or (idx.min() >= 0 and idx.max() < n))

    def _df(d, m, *, indices, **kw):
        d[...] = m.reshape(-1, order="F")[idx]
        return d

    def _dft(m, d, *, indices, **kw):
        m[...] = 0
        m.reshape(-1, order="F")[idx] = d
        return m
    return JopLn(df=_df, dft=_dft, dom=JetSpace(T, n), rng=JetSpace(T, idx.size), s={"indices": np.asarray(indices)})


def JopPointwise(T, n, fn="square", p=0.0):
    """d = phi(m);  δd = phi'(mo) .* δm  (self-adjoint)   (fixture JopBar,
    test/runtests.jl:20-25: ``δd .= 2 .* mₒ .* δm`` = ((2*mo)*δm))."""
    phi, dphi = PW_FUNCS[fn]

    def _f(d, m, **kw):
        d[...] = phi(m, p)
        return d

    def _df(dd, dm, *, mo, **kw):
        dd[...] = dphi(mo, p) * dm
        return dd
    spc = JetSpace(T, *((n,) if np.isscalar(n) else tuple(n)))
    return JopNl(f=_f, df=_df, dom=spc, rng=spc, s={"fn": fn, "p": p})


def JopStencil(T, n, kind="fdiff"):
    """1-D stencils on an n-vector (square n -> n so they chain with diagonals).

    fdiff :  d[i] = m[i+1] - m[i]  (i < n-1),  d[n-1] = 0
             adjoint  m[j] = (j>=1 ? d[j-1] : 0) - (j<=n-2 ? d[j] : 0)
    lap   :  d[i] = (m[i-1] - 2 m[i]) + m[i+1], out-of-range neighbours = 0 (self-adjoint)
    """
    n = int(n)

    def _fd(d, m, **kw):
        d[: n - 1] = m[1:] - m[: n - 1]
        d[n - 1] = 0
        return d

    def _fdt(m, d, **kw):
        left = np.zeros_like(d)
        left[1:] = d[: n - 1]
        right = d.copy()
        right[n - 1] = 0
        m[...] = left - right
        return m

    def _lap(d, m, **kw):
        left = np.zeros_like(m)
        left[1:] = m[: n - 1]
        right = np.zeros_like(m)
        right[: n - 1] = m[1:]
        d[...] = (left - m.dtype.type(2) * m) + right
        return d
    spc = JetSpace(T, n)
    if kind == "fdiff":
        return JopLn(df=_fd, dft=_fdt, dom=spc, rng=spc, s={"kind": kind})
    if kind == "lap":
        return JopLn(df=_lap, dom=spc, rng=spc, s={"kind": kind})
    raise ValueError(kind)


def JopDense(A):
    """d = A*m ; m = A'*d     (fixture JopBaz, test/runtests.jl:27-33; _matmul_df!,
    src/Jets.jl:573-574).  m may be a matrix of right-hand sides (cols x nrhs)."""
    def _df(d, m, *, A, **kw):
        d[...] = A @ m
        return d

    def _dft(m, d, *, A, **kw):
        m[...] = A.conj().T @ d
        return m
    return JopLn(df=_df, dft=_dft, dom=JetSpace(A.dtype, A.shape[1]),
                 rng=JetSpace(A.dtype, A.shape[0]), s={"A": A})
