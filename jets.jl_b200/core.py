"""Host-side mirror of the Jets public API for the operator-application path, over the C ABI.

Same names, argument meaning and error behaviour as the reference (ChevronETC/Jets.jl v1.4.1,
src/Jets.jl), so the parity tests read like the reference's own tests:

    JetSpace / JetBSpace            src/Jets.jl:40-68, 736-807
    zeros/ones/rand/randn/Array(R)  :105-108, :922-924      -> device-resident DeviceArray
    DeviceArray (Array/BlockArray)  :809-924                 getblock/setblock!/norm/dot/extrema/broadcast
    JopNl / JopLn / JopAdjoint      :194-228, mul! :390-392, * :399, adjoint :382-383
    jacobian / jacobian!            :364-375
    ∘ (compose, @), +, -, a*A       :569-576, :689-708, :1161-1164
    JopBlock / @blockop / JopZeroBlock / getblock / nblocks / isblockop   :926-1110
    dot_product_test / linearity_test / linearization_test / convert(Array, A)  :1174-1282

Julia's ``mul!(d, A, m)`` is ``mul_(d, A, m)``; ``A'`` is ``A.T`` / ``adjoint(A)``; ``A₂ ∘ A₁`` is
``compose(A2, A1)`` or ``A2 @ A1``.  Leaf operators come from the device primitive registry
(JopDiagonal, JopScale, JopPointwise, JopStencil, JopDense, JopZeroBlock) because arbitrary host
closures cannot run on the GPU.  All arithmetic happens in libjets_b200.so; numpy is used only to
move host data in and out.
"""
from __future__ import annotations

import ctypes as C
import itertools
import weakref
import math
from typing import Sequence

import numpy as np

from . import _lib as L
from ._lib import JetsError, lib, check

_DT = {np.dtype(np.float32): L.F32, np.dtype(np.float64): L.F64,
       np.dtype(np.complex64): L.C64, np.dtype(np.complex128): L.C128}
_NP = {v: k for k, v in _DT.items()}
_REAL = {np.dtype(np.complex64): np.dtype(np.float32), np.dtype(np.complex128): np.dtype(np.float64)}


def _dt(T):
    T = np.dtype(T)
    if T not in _DT:
        raise JetsError(3, f"eltype {T} is not supported on the device path (Float32/Float64/ComplexF32/ComplexF64)")
    return _DT[T]


def _iscomplex(T):
    return np.dtype(T).kind == "c"


# ------------------------------------------------------------------ spaces ----------------
class JetAbstractSpace:
    pass


class JetSpace(JetAbstractSpace):
    def __init__(self, T, *n):
        if len(n) == 1 and isinstance(n[0], (tuple, list)):
            n = tuple(n[0])
        self.T = np.dtype(T)
        self.n = tuple(int(k) for k in n)

    def __eq__(self, o):
        return isinstance(o, JetSpace) and self.T == o.T and self.n == o.n

    def __hash__(self):
        return hash((self.T, self.n))

    def __repr__(self):
        return f"JetSpace({self.T}, {self.n})"

    eltype = property(lambda s: s.T)
    ndims = property(lambda s: len(s.n))

    def size(self, i=None):
        return self.n if i is None else self.n[i - 1]

    def __len__(self):
        return int(np.prod(self.n, dtype=np.int64)) if self.n else 1

    def vec(self):
        return JetSpace(self.T, len(self))

    def similar(self, *dims):
        if len(dims) == 1 and isinstance(dims[0], (tuple, list)):
            dims = tuple(dims[0])
        return JetSpace(self.T, *dims)

    def _block_lens(self):
        return [len(self)]


class JetBSpace(JetAbstractSpace):
    """Block space with the reference's cumulative 1-based inclusive ranges (src/Jets.jl:742-748)."""

    def __init__(self, spaces: Sequence[JetAbstractSpace]):
        self.spaces = list(spaces)
        self.T = np.result_type(*[s.T for s in self.spaces])
        self.indices = []
        stop = 0
        for s in self.spaces:
            start = stop + 1
            stop = start + len(s) - 1
            self.indices.append((start, stop))

    def __eq__(self, o):
        return isinstance(o, JetBSpace) and self.spaces == o.spaces and self.indices == o.indices

    def __hash__(self):
        return hash(tuple(self.indices))

    def __repr__(self):
        return f"JetBSpace({self.spaces})"

    eltype = property(lambda s: s.T)
    ndims = property(lambda s: 1)

    def size(self, i=None):
        n = (self.indices[-1][1],)
        return n if i is None else n[i - 1]

    def __len__(self):
        return self.indices[-1][1]

    def vec(self):
        return self

    def similar(self, *dims):
        if len(dims) == 1 and isinstance(dims[0], (tuple, list)):
            dims = tuple(dims[0])
        return JetSpace(self.T, *dims)

    def _block_lens(self):
        return [len(s) for s in self.spaces]


class JetSSpace(JetAbstractSpace):
    """Symmetric space (src/Jets.jl:408-441): logical size ``n``, stored parent of size ``M``;
    an index beyond the parent in some dimension reads ``conj(parent[map(I)])`` (:455-462).
    ``map`` takes and returns a 1-based index tuple, as the reference's ``map(I)`` does with a
    CartesianIndex."""

    def __init__(self, T, n, M, map):
        self.T = np.dtype(T)
        self.n = tuple(int(k) for k in n)
        self.M = tuple(int(k) for k in M)
        self.map = map
        self._weights = None   # device vector: logical entries represented by each stored element

    def __eq__(self, o):
        return isinstance(o, JetSSpace) and self.T == o.T and self.n == o.n and self.M == o.M and self.map is o.map

    def __hash__(self):
        return hash((self.T, self.n, self.M))

    def __repr__(self):
        return f"JetSSpace({self.T}, {self.n}, {self.M})"

    eltype = property(lambda s: s.T)
    ndims = property(lambda s: len(s.n))

    def size(self, i=None):
        return self.n if i is None else self.n[i - 1]

    def __len__(self):
        return int(np.prod(self.n, dtype=np.int64))

    def similar(self, *dims):
        if len(dims) == 1 and isinstance(dims[0], (tuple, list)):
            dims = tuple(dims[0])
        return JetSSpace(self.T, dims, self.M, self.map)

    def _block_lens(self):   # what is stored: the parent
        return [int(np.prod(self.M, dtype=np.int64))]

    def _beyond(self, I):
        return any(I[d] > self.M[d] for d in range(len(self.n)))

    def _parent_linear(self, I):
        """0-based column-major offset into the parent of the stored element behind logical index I
        (1-based tuple), and whether it enters conjugated."""
        cj = self._beyond(I)
        J = tuple(self.map(tuple(I))) if cj else tuple(I)
        off, stride = 0, 1
        for d, j in enumerate(J):
            off += (int(j) - 1) * stride
            stride *= self.M[d]
        return off, cj

    def multiplicity(self):
        """Host table, one entry per stored element: 1 + the number of mirrored logical positions
        that map onto it -- the weights with which norm() over the logical array (generic
        AbstractArray iteration through getindex :455-462) sees each stored value."""
        w = np.ones(self._block_lens()[0], dtype=np.float64)
        for I in itertools.product(*[range(1, k + 1) for k in reversed(self.n)]):
            I = tuple(reversed(I))
            if self._beyond(I):
                w[self._parent_linear(I)[0]] += 1.0
        return w


def symspace():
    return None


def indices(R, iblock):
    return R.indices[iblock - 1]


# ------------------------------------------------------------------ device arrays ---------
class DeviceArray:
    """A device-resident vector in a JetSpace (plain array) or JetBSpace (BlockArray,
    src/Jets.jl:809-812).  Storage is one flat buffer; block i occupies the reference's index
    range ``indices(R, i)``."""

    __array_priority__ = 1000
    __array_ufunc__ = None

    def __init__(self, handle, space_, owner=None):
        self._h = C.c_void_p(handle) if not isinstance(handle, C.c_void_p) else handle
        self.space = space_
        self._owner = owner  # keeps a wrapped torch tensor / parent alive

    def __del__(self):
        try:
            if self._h and lib is not None:
                lib.jets_buf_destroy(self._h)
        except Exception:
            pass

    # --- metadata
    @property
    def dtype(self):
        return self.space.T

    def __len__(self):
        return len(self.space)

    @property
    def shape(self):
        return self.space.size()

    @property
    def isblock(self):
        return isinstance(self.space, JetBSpace)

    @property
    def indices(self):
        if self.isblock:
            return self.space.indices
        return [(1, len(self))]

    def block_range(self, iblock):
        """indices(x, i): read back from the device block table (1-based inclusive)."""
        a, b = C.c_int64(), C.c_int64()
        check(lib.jets_buf_block_range(self._h, iblock - 1, C.byref(a), C.byref(b)))
        return a.value, b.value

    @property
    def devptr(self):
        return lib.jets_buf_devptr(self._h)

    # --- host transfer
    def to_host(self):
        """convert(Array, x) (src/Jets.jl:862-868): flat host vector for block arrays, shaped
        (column-major) array for plain spaces."""
        if isinstance(self.space, JetSSpace):
            return self.full()
        out = np.empty(len(self), dtype=self.dtype)
        check(lib.jets_buf_download(self._h, -1, out.ctypes.data_as(C.c_void_p), out.size))
        if self.isblock:
            return out
        return out.reshape(self.space.n, order="F")

    # --- SymmetricArray interface (src/Jets.jl:443-484)
    @property
    def issymmetric(self):
        return isinstance(self.space, JetSSpace)

    def parent(self):
        """parent(A::SymmetricArray) (:449): the stored part, as a host array of size M."""
        R = self.space
        if not isinstance(R, JetSSpace):
            return self.to_host()
        out = np.empty(R._block_lens()[0], dtype=self.dtype)
        check(lib.jets_buf_download(self._h, -1, out.ctypes.data_as(C.c_void_p), out.size))
        return out.reshape(R.M, order="F")

    A = property(lambda self: self.parent())

    def full(self):
        """The logical n-sized array, mirrored entries filled in through getindex (:455-462)."""
        R = self.space
        P = self.parent().reshape(-1, order="F")
        out = np.empty(R.n, dtype=self.dtype)
        for I in itertools.product(*[range(1, k + 1) for k in R.n]):
            off, cj = R._parent_linear(I)
            out[tuple(i - 1 for i in I)] = np.conj(P[off]) if cj else P[off]
        return out

    def _elem_offset(self, I):
        """(0-based storage offset, conjugated?) of a 1-based index: tuple (Cartesian) or int (linear,
        column-major over the logical size, :465-468)."""
        R = self.space
        if isinstance(R, JetSSpace):
            if not isinstance(I, tuple):
                I = tuple(int(k) + 1 for k in np.unravel_index(int(I) - 1, R.n, order="F"))
            return R._parent_linear(I)
        if isinstance(I, tuple):
            I = int(np.ravel_multi_index(tuple(int(k) - 1 for k in I), R.n, order="F")) + 1
        return int(I) - 1, False

    def __getitem__(self, I):
        """x[i] / x[i1,i2,...] with the reference's 1-based indices (scalar getindex, :819-825, :455-468)."""
        off, cj = self._elem_offset(I)
        v = np.empty(1, dtype=self.dtype)
        check(lib.jets_buf_read(self._h, off, v.ctypes.data_as(C.c_void_p), 1))
        return np.conj(v[0]) if cj else v[0]

    def __setitem__(self, I, val):
        """x[i] = v (scalar setindex!, :826-832; mirrored positions store conj(v), :470-478)."""
        off, cj = self._elem_offset(I)
        v = np.array([np.conj(val) if cj else val], dtype=self.dtype)
        check(lib.jets_buf_write(self._h, off, v.ctypes.data_as(C.c_void_p), 1))

    def from_host(self, x):
        x = np.asarray(x)
        if isinstance(self.space, JetSSpace) and x.shape == self.space.n and self.space.n != self.space.M:
            x = x[tuple(slice(0, m) for m in self.space.M)]   # keep the stored part of a full array
        if x.ndim > 1:
            x = x.reshape(-1, order="F")
        x = np.ascontiguousarray(x, dtype=self.dtype)
        check(lib.jets_buf_upload(self._h, -1, x.ctypes.data_as(C.c_void_p), x.size))
        return self

    # --- BlockArray interface
    def getblock(self, iblock):
        """getblock(x, i): a view sharing memory (src/Jets.jl:914)."""
        if not self.isblock:
            return self
        h = C.c_void_p()
        check(lib.jets_buf_view(self._h, iblock - 1, 1, C.byref(h)))
        return DeviceArray(h, self.space.spaces[iblock - 1], owner=self)

    def setblock_(self, iblock, xblock):
        blk = self.getblock(iblock)
        if isinstance(xblock, DeviceArray):
            check(lib.jets_buf_copy(blk._h, xblock._h))
        elif np.isscalar(xblock):
            blk.fill_(xblock)
        else:
            blk.from_host(xblock)
        return blk

    def fill_(self, a):
        a = complex(a)
        check(lib.jets_buf_fill_c(self._h, a.real, a.imag))
        return self

    def assign(self, src):
        """``self .= src``."""
        if isinstance(src, DeviceArray):
            check(lib.jets_buf_copy(self._h, src._h))
        elif np.isscalar(src):
            self.fill_(src)
        else:
            self.from_host(src)
        return self

    def copy(self):
        return similar(self).assign(self)

    # --- broadcast arithmetic (src/Jets.jl:889-911): one fused device pass per expression
    def _lin(self, terms):
        out = similar(self)
        lincomb_(out, terms)
        return out

    def __add__(self, o):
        return self._lin([(1.0, self), (1.0, _as_dev(o, self))])

    __radd__ = __add__

    def __sub__(self, o):
        return self._lin([(1.0, self), (-1.0, _as_dev(o, self))])

    def __rsub__(self, o):
        return self._lin([(1.0, _as_dev(o, self)), (-1.0, self)])

    def __neg__(self):
        return self._lin([(-1.0, self)])

    def __mul__(self, o):
        if np.isscalar(o):
            return self._lin([(o, self)])
        out = similar(self)
        check(lib.jets_hadamard(out._h, self._h, _as_dev(o, self)._h))
        return out

    __rmul__ = __mul__

    def __truediv__(self, o):
        if np.isscalar(o):
            return self._lin([(1.0 / o, self)])
        return NotImplemented

    def __abs__(self):
        return abs_(self)


def _as_dev(o, like: DeviceArray) -> DeviceArray:
    if isinstance(o, DeviceArray):
        return o
    return similar(like).from_host(np.broadcast_to(np.asarray(o, dtype=like.dtype), (len(like),))
                                   if np.ndim(o) == 0 else o)


def _alloc(R) -> DeviceArray:
    L.ensure_init()
    lens = R._block_lens()
    arr = (C.c_int64 * len(lens))(*lens)
    h = C.c_void_p()
    check(lib.jets_buf_create(_dt(R.T), len(lens), arr, C.byref(h)))
    return DeviceArray(h, R)


_seed = itertools.count(0x5EED)


def zeros(R):
    return _alloc(R)


def Array(R):
    return _alloc(R)


def ones(R):
    return _alloc(R).fill_(1.0)


def rand(R, seed=None):
    x = _alloc(R)
    check(lib.jets_buf_rand(x._h, next(_seed) if seed is None else int(seed), 0, 0))
    return x


def randn(R, seed=None):
    x = _alloc(R)
    check(lib.jets_buf_rand(x._h, next(_seed) if seed is None else int(seed), 0, 1))
    return x


def similar(x: DeviceArray):
    return _alloc(x.space)


def to_device(x, R=None) -> DeviceArray:
    """Host array -> device array in space R (default: space(x))."""
    if isinstance(x, DeviceArray):
        return x
    x = np.asarray(x)
    R = JetSpace(x.dtype, *x.shape) if R is None else R
    return _alloc(R).from_host(x)


def wrap_torch(t, R=None) -> DeviceArray:
    """Zero-copy view of a CUDA torch tensor (jets_buf_wrap): no guard padding, so fused applies
    on it use the guarded-load engine rather than TMA."""
    L.ensure_init()
    assert t.is_cuda and t.is_contiguous()
    T = {4: np.float32, 8: np.float64}[t.element_size()]
    R = JetSpace(T, t.numel()) if R is None else R
    lens = R._block_lens()
    arr = (C.c_int64 * len(lens))(*lens)
    h = C.c_void_p()
    check(lib.jets_buf_wrap(_dt(R.T), C.c_void_p(t.data_ptr()), len(lens), arr, C.byref(h)))
    return DeviceArray(h, R, owner=t)


def reshape(x: DeviceArray, R) -> DeviceArray:
    """reshape(x, R) (src/Jets.jl:38, :1112-1118): same memory, new space."""
    if len(x) != len(R):
        raise JetsError(2, "dimension mismatch, unable to reshape block array")
    lens = R._block_lens()
    arr = (C.c_int64 * len(lens))(*lens)
    h = C.c_void_p()
    check(lib.jets_buf_reshape(x._h, len(lens), arr, C.byref(h)))
    return DeviceArray(h, R, owner=x)


def space(x, iblock=None):
    if isinstance(x, JetBSpace):
        return x.spaces[iblock - 1]
    if isinstance(x, DeviceArray):
        return x.space
    x = np.asarray(x)
    return JetSpace(x.dtype, *x.shape)


def getblock(x, i, j=None):
    if j is not None:
        return _getblock_op(x, i, j)
    return x.getblock(i) if isinstance(x, DeviceArray) else x


def getblock_(x, iblock, xblock):
    """getblock!(x, i, xblock): copy block i into xblock (host array or device array)."""
    b = x.getblock(iblock)
    if isinstance(xblock, DeviceArray):
        return xblock.assign(b)
    xblock[...] = b.to_host().reshape(xblock.shape, order="F")
    return xblock


def setblock_(x, iblock, xblock):
    return x.setblock_(iblock, xblock)


def fill_(x, a):
    return x.fill_(a)


def to_array(x):
    return x.to_host() if isinstance(x, DeviceArray) else np.asarray(x)


def lincomb_(out: DeviceArray, terms):
    """out .= c1.*x1 .+ c2.*x2 ...  (<= 4 terms per pass; longer sums chain through out)."""
    terms = list(terms)
    first = True
    while terms:
        chunk = terms[: (4 if first else 3)]
        terms = terms[len(chunk):]
        if not first:
            chunk = [(1.0, out)] + chunk
        n = len(chunk)
        xs = (C.c_void_p * n)(*[x._h for _, x in chunk])
        if any(isinstance(c, complex) or np.iscomplexobj(c) for c, _ in chunk):
            flat = []
            for c, _ in chunk:
                c = complex(c)
                flat += [c.real, c.imag]
            check(lib.jets_lincomb_c(out._h, n, (C.c_double * (2 * n))(*flat), xs))
        else:
            cs = (C.c_double * n)(*[float(c) for c, _ in chunk])
            check(lib.jets_lincomb(out._h, n, cs, xs))
        first = False
    return out


def abs_(x: DeviceArray) -> DeviceArray:
    """abs.(x) for a complex vector: a real vector on the same block structure (test/runtests.jl:545-547)."""
    R = x.space
    T = _REAL[np.dtype(x.dtype)]
    if isinstance(R, JetBSpace):
        Rr = JetBSpace([JetSpace(T, *s.n) for s in R.spaces])
    elif isinstance(R, JetSSpace):
        Rr = JetSpace(T, *R.M)
    else:
        Rr = JetSpace(T, *R.n)
    out = _alloc(Rr)
    check(lib.jets_abs(out._h, x._h))
    return out


def hadamard_(out, x, y):
    check(lib.jets_hadamard(out._h, x._h, y._h))
    return out


def dot(x: DeviceArray, y: DeviceArray):
    """dot(x, y) (src/Jets.jl:850-856); conj on the first argument for complex eltypes."""
    if _iscomplex(x.dtype):
        r = (C.c_double * 2)()
        check(lib.jets_dot_c(x._h, y._h, r))
        return x.dtype.type(complex(r[0], r[1]))
    r = C.c_double()
    check(lib.jets_dot(x._h, y._h, C.byref(r)))
    return x.dtype.type(r.value)


def _sym_weights(R: JetSSpace) -> DeviceArray:
    if R._weights is None:
        w = R.multiplicity()
        R._weights = _alloc(JetSpace(np.float64, w.size)).from_host(w)
    return R._weights


def norm(x: DeviceArray, p=2):
    r = C.c_double()
    if isinstance(x.space, JetSSpace):
        # norm over the logical array: every stored element counts once per position it stands for
        check(lib.jets_norm_weighted(x._h, _sym_weights(x.space)._h, float(p), C.byref(r)))
    else:
        check(lib.jets_norm(x._h, float(p), C.byref(r)))
    return _REAL.get(np.dtype(x.dtype), np.dtype(x.dtype)).type(r.value)


def extrema(x: DeviceArray):
    a, b = C.c_double(), C.c_double()
    check(lib.jets_extrema(x._h, C.byref(a), C.byref(b)))
    return x.dtype.type(a.value), x.dtype.type(b.value)


def nblocks(x, i=None):
    if isinstance(x, JetBSpace):
        return len(x.spaces)
    if isinstance(x, JetAbstractSpace):
        return 1
    if isinstance(x, DeviceArray):
        return nblocks(x.space)
    nb = (nblocks(range_(x)), nblocks(domain(x)))
    return nb if i is None else nb[i - 1]


def sync():
    check(lib.jets_sync())


# ------------------------------------------------------------------ operators -------------
class _Handle:
    """Owns one reference to a jets_op."""

    def __init__(self, h):
        self.h = h if isinstance(h, C.c_void_p) else C.c_void_p(h)

    def __del__(self):
        try:
            if self.h and lib is not None:
                lib.jets_op_destroy(self.h)
        except Exception:
            pass


def _newop(fn, *args):
    L.ensure_init()
    h = C.c_void_p()
    check(fn(*args, C.byref(h)))
    return _Handle(h)


class Jop:
    """Common behaviour of JopNl / JopLn / JopAdjoint.  ``_h`` is the operator tree handle that is
    applied for this wrapper; ``_mode`` the mul! dispatch (src/Jets.jl:390-392)."""

    __array_priority__ = 2000
    __array_ufunc__ = None
    _mode = L.MODE_DF

    def __init__(self, handle: _Handle, dom, rng, meta=None):
        self._h = handle
        self.dom, self.rng = dom, rng
        self.meta = meta or {}

    # A*m (src/Jets.jl:399).  Host arrays take the end-to-end path: upload, apply, download.
    def __mul__(self, m):
        if isinstance(m, DeviceArray):
            return mul_(zeros(range_(self)), self, m)
        if isinstance(m, np.ndarray):
            md = to_device(m, domain(self))
            return mul_(zeros(range_(self)), self, md).to_host()
        return NotImplemented

    def __rmul__(self, a):
        if np.isscalar(a):
            return scalar_mul(a, self)
        return NotImplemented

    def __matmul__(self, o):
        return compose(self, o)

    def __rmatmul__(self, o):
        return compose(o, self)

    def __add__(self, o):
        return op_sum(self, o, +1)

    def __radd__(self, o):
        return op_sum(o, self, +1)

    def __sub__(self, o):
        return op_sum(self, o, -1)

    def __rsub__(self, o):
        return op_sum(o, self, -1)

    @property
    def T(self):
        return adjoint(self)

    def close(self):
        """Base.close(A) (src/Jets.jl:290,1120): release the device resources of this handle."""
        self._h = None


class JopNl(Jop):
    _mode = L.MODE_F


class JopLn(Jop):
    _mode = L.MODE_DF


class JopAdjoint(Jop):
    _mode = L.MODE_DF  # the handle is the adjoint view; its DF is the parent's df'!

    def __init__(self, op: Jop, handle: _Handle):
        super().__init__(handle, op.rng, op.dom, op.meta)
        self.op = op


def _is_lin(A):
    return isinstance(A, (JopLn, JopAdjoint))


class Jet:
    """The jet behind an operator (src/Jets.jl:133-142), device edition.  ``dom``, ``rng``, the linearization point and the
    state are the reference's fields; the three mappings f!/df!/df'! live in libjets_b200 behind the operator handle --
    the primitive registry (``jets_op_*``) takes the place of plugin closures on the device -- so a Jet is obtained from
    an operator (``jet(A)``) and cannot be built from host closures: ``Jet(dom=..., f=...)`` raises
    JETS_ERR_UNSUPPORTED, the device path's ``error("not implemented")`` (:131).  ``jet(F)`` and ``jet(JopLn(F))`` are
    the SAME object, as in the reference (:364-366), and every accessor that takes a Jop takes a Jet:
    ``domain / range_ / shape / state / state_ / point / point_ / perfstat / close``."""

    def __init__(self, *args, **kw):
        raise JetsError(5, "Jet(dom, rng, f!, df!, df'!) takes host closures, which cannot run on the device: build the "
                           "operator from the library's leaves (JopDiagonal, JopPointwise, JopStencil, JopDense, "
                           "JopRestriction, JopZeroBlock) and the combinators, and use jet(A) for its jet")

    @classmethod
    def _of(cls, A):
        j = object.__new__(cls)
        j._h, j.dom, j.rng, j.meta = A._h, A.dom, A.rng, A.meta
        return j

    _mode = L.MODE_F

    @property
    def mo(self):
        return self.meta.get("mo")

    @property
    def s(self):
        return self.meta

    def close(self):      # the handle belongs to the operator(s) of this jet
        return None

    def __repr__(self):
        return f"Jet({self.dom} -> {self.rng})"


def jet(A):
    """jet(A) (src/Jets.jl:236-238): the jet of an operator; the adjoint's jet is its parent's."""
    if isinstance(A, Jet):
        return A
    if isinstance(A, JopAdjoint):
        return jet(A.op)
    j = _JETS.get(id(A.meta))        # operators that share a state dictionary share a jet (jacobian! :364-366)
    if j is None or j.meta is not A.meta:
        j = _JETS[id(A.meta)] = Jet._of(A)
    return j


_JETS = weakref.WeakValueDictionary()


def domain(A):
    if isinstance(A, np.ndarray):
        return JetSpace(A.dtype, A.shape[1])
    return A.dom


def range_(A):
    if isinstance(A, np.ndarray):
        return JetSpace(A.dtype, A.shape[0])
    return A.rng


def eltype(A):
    return np.result_type(domain(A).T, range_(A).T)


def shape(A, i=None):
    s = (range_(A).size(), domain(A).size())
    return s if i is None else s[0 if i == 1 else 1]


def size(A, i=None):
    s = (len(range_(A)), len(domain(A)))
    return s if i is None else s[0 if i == 1 else 1]


def state(A, key=None):
    """state(A[, key]) (src/Jets.jl:264-265, :313-314).  For a composition a missing key is looked up
    in the operands and must be unambiguous (:607-623)."""
    if key is None:
        return A.meta
    if key in A.meta:
        return A.meta[key]
    if A.meta.get("kind") == "compose":
        hits = [o for o in A.meta["ops"] if key in o.meta]
        if not hits:
            raise KeyError(f"key {key} does not exist in the state of the composite operator")
        if len(hits) > 1:
            raise KeyError(f"ambiguous: key {key} exists in more than one operator in the composition")
        return state(hits[0], key)
    raise KeyError(key)


def state_(A, s):
    """state!(A, s) (src/Jets.jl:272, :315): merge ``s`` into the operator's state.  State that lives on
    the device (a diagonal, a matrix) is updated IN the buffer the kernels read, so the new values take
    effect on the next mul! exactly as a Jets closure would see its new keyword arguments."""
    for k, v in dict(s).items():
        cur = A.meta.get(k)
        if isinstance(cur, DeviceArray):
            cur.assign(v)
        elif k in _BAKED_STATE and k in A.meta:
            # constants baked into the device operator when it was built (a Jets closure would see the new keyword
            # argument; the kernels would silently keep the old value)
            raise JetsError(5, f"state!: '{k}' is compiled into the device operator; build a new operator instead")
        else:
            A.meta[k] = v
    return A


_BAKED_STATE = ("a", "p", "fn", "kind", "indices", "zero")


def perfstat(A):
    """perfstat(A) (src/Jets.jl:281, :316; composite :597-605): the engine(s) the planner chose for this
    operator and the number of kernel launches of one apply."""
    if isinstance(A, JopAdjoint):
        return perfstat(A.op)
    return plan_info(A)


def close(A):
    """Base.close(A) (src/Jets.jl:290, :317; composite :591-595): leaves return False, combinators close
    their operands and return None.  The device handle is released when its last holder goes away."""
    if isinstance(A, JopAdjoint):
        return close(A.op)
    kids = A.meta.get("ops")
    A.close()
    if kids is None:
        return False
    for o in (kids.ravel() if isinstance(kids, np.ndarray) else kids):
        if isinstance(o, Jop):
            close(o)
    return None


def _lin_handle(A: Jop) -> _Handle:
    """Handle of JopLn(A) as a child: a linear view of the same jet (src/Jets.jl:209-224)."""
    if _is_lin(A):
        return A._h
    return _newop(lib.jets_op_as_linear, A._h.h)


def adjoint(A):
    if isinstance(A, np.ndarray):
        return A.conj().T
    if isinstance(A, JopAdjoint):
        return A.op
    if isinstance(A, JopLn):
        return JopAdjoint(A, _newop(lib.jets_op_adjoint, A._h.h))
    raise JetsError(6, "adjoint is defined for JopLn/JopAdjoint only (src/Jets.jl:382-383)")


def mul_(d: DeviceArray, A: Jop, m: DeviceArray, accumulate: bool = False):
    """mul!(d, A, m) (src/Jets.jl:390-392).  ``accumulate=True`` reproduces reference quirk Q1:
    a forward JopBlock apply with more than one block column adds into ``d``
    (src/Jets.jl:1001,1024); the default overwrites ``d``, which is what the reference computes
    whenever ``d`` was zero-initialised, i.e. for every ``A*m``."""
    check(lib.jets_apply(A._h.h, A._mode, d._h, m._h, 1 if accumulate else 0))
    return d


def point_(A: Jop, mo: DeviceArray):
    check(lib.jets_op_set_point(A._h.h, mo._h))
    A.meta["mo"] = mo
    return A


def point(A: Jop):
    return A.meta.get("mo")


def jacobian_(F, mo: DeviceArray):
    """jacobian!(F, mo) (src/Jets.jl:364-366): shares and mutates the underlying jet."""
    if _is_lin(F) or isinstance(F, np.ndarray):
        return F
    point_(F, mo)
    return JopLn(_lin_handle(F), F.dom, F.rng, F.meta)


def copy(A, copymo=True):
    """copy(A[, copymₒ]) (src/Jets.jl:230-233): a new jet of the same shape.  The operator's state buffers
    are immutable on the device and therefore shared; the linearization point is held
    by reference as in ``copy(A, false)`` -- with ``copymo`` a private snapshot is taken at the next
    ``point_``/``jacobian`` anyway, which is when the reference's copy becomes observable."""
    if isinstance(A, JopAdjoint):
        return adjoint(copy(A.op, copymo))
    h = _newop(lib.jets_op_clone, A._h.h)
    return type(A)(h, A.dom, A.rng, dict(A.meta))


def jacobian(F, mo: DeviceArray):
    """jacobian(F, mo) (src/Jets.jl:374): new jet with a private snapshot of mo; the (immutable)
    state buffers are shared instead of deep-copied."""
    if isinstance(F, np.ndarray):
        return F.copy()
    if isinstance(F, JopAdjoint):
        return adjoint(jacobian(F.op, mo))
    h = _newop(lib.jets_op_jacobian, F._h.h, mo._h)
    meta = dict(F.meta)
    meta["mo"] = mo
    return JopLn(h, F.dom, F.rng, meta)


# ---- primitive registry -------------------------------------------------------------------
def JopDiagonal(w):
    """d = w .* m (fixture JopFoo, test/runtests.jl:3-8).  w: host array or DeviceArray."""
    wd = to_device(w)
    sp = wd.space if isinstance(wd.space, JetSpace) else JetSpace(wd.dtype, len(wd))
    return JopLn(_newop(lib.jets_op_diag, wd._h), sp, sp, {"diagonal": wd})


def JopScale(T, n, a):
    sp = JetSpace(T, *((n,) if np.isscalar(n) else tuple(n)))
    a_ = complex(a)
    return JopLn(_newop(lib.jets_op_scale_c, _dt(T), len(sp), a_.real, a_.imag), sp, sp, {"a": a})


def JopPointwise(T, n, fn="square", p=0.0):
    """d = phi(m); Jacobian phi'(mo) .* dm (fixture JopBar, test/runtests.jl:20-25)."""
    sp = JetSpace(T, *((n,) if np.isscalar(n) else tuple(n)))
    return JopNl(_newop(lib.jets_op_pointwise, _dt(T), len(sp), L.PW[fn], float(p)), sp, sp,
                 {"fn": fn, "p": p})


def JopStencil(T, n, kind="fdiff"):
    sp = JetSpace(T, int(n))
    return JopLn(_newop(lib.jets_op_stencil, _dt(T), int(n), L.STENCIL[kind]), sp, sp, {"kind": kind})


def JopRestriction(T, n, indices):
    """d = m[indices]; adjoint m .= 0, m[indices] = d (JetPack-style restriction; ``indices`` 1-based and
    unique, as a Julia caller passes them)."""
    idx = np.ascontiguousarray(np.asarray(indices, dtype=np.int64).reshape(-1) - 1)
    L.ensure_init()
    h = C.c_void_p()
    check(lib.jets_op_restrict(_dt(T), int(n), idx.size, idx.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(h)))
    return JopLn(_Handle(h), JetSpace(T, int(n)), JetSpace(T, idx.size), {"indices": np.asarray(indices)})


def JopDense(A, nrhs=1):
    """d = A*m, m = A'*d (fixture JopBaz test/runtests.jl:27-33; matrix interop src/Jets.jl:573-576).
    A: host (rows x cols) array or a DeviceArray in JetSpace(T, rows, cols) (column-major)."""
    if isinstance(A, DeviceArray):
        Ad = A
        rows, cols = A.space.n
    else:
        A = np.asarray(A)
        rows, cols = A.shape
        Ad = to_device(np.asfortranarray(A), JetSpace(A.dtype, rows, cols))
    T = Ad.dtype
    dom = JetSpace(T, cols) if nrhs == 1 else JetSpace(T, cols, nrhs)
    rng = JetSpace(T, rows) if nrhs == 1 else JetSpace(T, rows, nrhs)
    return JopLn(_newop(lib.jets_op_dense, Ad._h, rows, cols, nrhs), dom, rng, {"A": Ad})


def JopZeroBlock(dom, rng):
    return JopLn(_newop(lib.jets_op_zero, _dt(dom.T), len(dom), len(rng)), dom, rng, {"zero": True})


def iszero(A):
    return bool(lib.jets_op_is_zero(A._h.h))


def isblockop(A):
    return bool(lib.jets_op_is_block(A._h.h))


# ---- combinators ----------------------------------------------------------------------------
def _wrap_matrix(A):
    return JopDense(A)


def compose(A2, A1):
    """A2 ∘ A1 (src/Jets.jl:569-576)."""
    if isinstance(A2, np.ndarray) and isinstance(A1, np.ndarray):
        return A2 @ A1
    if isinstance(A1, np.ndarray):
        A1 = _wrap_matrix(A1)
    if isinstance(A2, np.ndarray):
        A2 = _wrap_matrix(A2)
    hs = (C.c_void_p * 2)(A2._h.h, A1._h.h)
    h = _newop(lib.jets_op_compose, 2, hs)
    meta = {"kind": "compose", "ops": _comp_ops(A2) + _comp_ops(A1)}
    cls = JopLn if (_is_lin(A2) and _is_lin(A1)) else JopNl
    return cls(h, domain(A1), range_(A2), meta)


def _comp_ops(A):
    """jops_comp (src/Jets.jl:542-550): flattened operand tuple kept on the host for state(A).ops."""
    if isinstance(A, JopAdjoint):
        if A.op.meta.get("kind") == "compose":
            return tuple(adjoint(o) if _is_lin(o) else o for o in reversed(A.op.meta["ops"]))
        return (A,)
    if A.meta.get("kind") == "compose":
        return tuple(A.meta["ops"])
    return (A,)


def op_sum(A2, A1, sign):
    """A2 ± A1 (src/Jets.jl:689-708)."""
    if isinstance(A1, np.ndarray):
        A1 = _wrap_matrix(A1)
    if isinstance(A2, np.ndarray):
        A2 = _wrap_matrix(A2)
    hs = (C.c_void_p * 2)(A2._h.h, A1._h.h)
    sg = (C.c_int32 * 2)(1, sign)
    h = _newop(lib.jets_op_sum, 2, hs, sg)
    s2 = A2.meta["sgns"] if A2.meta.get("kind") == "sum" else (1,)
    s1 = A1.meta["sgns"] if A1.meta.get("kind") == "sum" else (1,)
    meta = {"kind": "sum", "sgns": tuple(s2) + tuple(s * sign for s in s1)}
    cls = JopLn if (_is_lin(A2) and _is_lin(A1)) else JopNl
    return cls(h, domain(A2), range_(A2), meta)


def scalar_mul(a, A):
    """a*A (src/Jets.jl:1161-1164)."""
    a_ = complex(a)
    h = _newop(lib.jets_op_scalar_mul_c, a_.real, a_.imag, A._h.h)
    cls = JopLn if _is_lin(A) else JopNl
    return cls(h, domain(A), range_(A), {})


def _as_op_matrix(ops):
    if isinstance(ops, np.ndarray) and ops.dtype == object:
        return ops.reshape(-1, 1) if ops.ndim == 1 else ops
    if len(ops) and isinstance(ops[0], (list, tuple)):
        a = np.empty((len(ops), len(ops[0])), dtype=object)
        for i, row in enumerate(ops):
            for k, o in enumerate(row):
                a[i, k] = o
        return a
    a = np.empty((len(ops), 1), dtype=object)
    for i, o in enumerate(ops):
        a[i, 0] = o
    return a


def JopBlock(ops, dadom=False, **kw):
    """JopBlock / @blockop (src/Jets.jl:926-986).  ``ops``: list of rows, 2-D object array, or a
    vector of operators (= one block column, :933)."""
    arr = _as_op_matrix(ops)
    nr, nc = arr.shape
    flat = [arr[r, c]._h.h for c in range(nc) for r in range(nr)]  # column-major like Julia
    hs = (C.c_void_p * len(flat))(*flat)
    h = _newop(lib.jets_op_block, nr, nc, hs, 1 if dadom else 0)
    dom = (domain(arr[0, 0]) if (nc == 1 and not dadom)
           else JetBSpace([domain(arr[0, c]) for c in range(nc)]))
    rng = JetBSpace([range_(arr[r, 0]) for r in range(nr)])
    lin = all(_is_lin(o) for o in arr.reshape(-1))
    meta = {"ops": arr, "kind": "block"}
    meta.update(kw)
    return (JopLn if lin else JopNl)(h, dom, rng, meta)


blockop = JopBlock


def _getblock_op(A, i, j):
    """getblock(A, i, j) (src/Jets.jl:1085-1110), 1-based."""
    h = _newop(lib.jets_op_getblock, A._h.h, i - 1, j - 1)
    lin = bool(lib.jets_op_is_linear(h.h))
    nd = lib.jets_op_nblocks(h.h, 2)
    T = _NP[lib.jets_op_dtype(h.h)]

    def _sp(which):
        n = lib.jets_op_nblocks(h.h, which)
        lens = []
        for b in range(n):
            v = C.c_int64()
            check(lib.jets_op_block_len(h.h, which, b, C.byref(v)))
            lens.append(v.value)
        return JetSpace(T, lens[0]) if n == 1 else JetBSpace([JetSpace(T, l) for l in lens])
    del nd
    dom, rng = _sp(2), _sp(1)
    if isinstance(A, JopAdjoint):
        inner = _getblock_op(A.op, j, i)
        return JopAdjoint(inner if _is_lin(inner) else JopLn(_lin_handle(inner), inner.dom, inner.rng, inner.meta), h)
    return (JopLn if lin else JopNl)(h, dom, rng, {})


# ---- utilities --------------------------------------------------------------------------------
def to_matrix(A):
    """convert(Array, A::Jop) (src/Jets.jl:1174-1185): column by column on the device."""
    m = zeros(domain(A))
    d = zeros(range_(A))
    nr, nc = size(A)
    B = np.zeros((nr, nc), dtype=eltype(A))
    e = np.zeros(nc, dtype=domain(A).T)
    for icol in range(nc):
        e[:] = 0
        e[icol] = 1
        m.from_host(e)
        d.fill_(0)
        B[:, icol] = mul_(d, A, m).to_host().reshape(-1, order="F")
    return B


def dot_product_test(op, m, d, mmask=None, dmask=None):
    """(src/Jets.jl:1211-1226) returns (lhs, rhs) = (<mmask.*m, A'(dmask.*d)>, <A(mmask.*m), dmask.*d>)."""
    if not isinstance(op, JopLn):
        raise JetsError(6, "dot_product_test accepts JopLn only")
    mmask = ones(domain(op)) if mmask is None else mmask
    dmask = ones(range_(op)) if dmask is None else dmask
    mm = mmask * m
    dd = dmask * d
    ds = op * mm
    ms = adjoint(op) * dd
    return dot(mm, ms), dot(ds, dd)


def linearity_test(A, m1=None, m2=None):
    """(src/Jets.jl:1276-1282); default vectors are -2*rand (reference quirk Q7)."""
    m1 = rand(domain(A)) * -2.0 if m1 is None else m1
    m2 = rand(domain(A)) * -2.0 if m2 is None else m2
    lhs = A * (m1 + m2)
    rhs = (A * m1) + (A * m2)
    return lhs, rhs


def linearization_test(F, mo, mu=(1.0, 0.5, 0.25, 0.125, 0.0625, 0.03125), dm=None, seed=None):
    """(src/Jets.jl:1235-1266) Taylor remainder ratios, evaluated on the device."""
    dm = (rand(domain(F), seed) * 2.0 - _as_dev(1.0, mo)) if dm is None else dm
    Fo = F * mo
    Jo = jacobian_(F, mo)
    Jodm = Jo * dm
    mu = sorted(mu, reverse=True)
    phi = np.zeros(len(mu))
    muobs, muexp = np.zeros(len(mu) - 1), np.zeros(len(mu) - 1)
    for i, u in enumerate(mu):
        d_lin = Fo + Jodm * u
        d_non = F * (mo + dm * u)
        phi[i] = float(norm(d_non - d_lin))
        if i > 0:
            muobs[i - 1] = phi[i - 1] / phi[i]
            muexp[i - 1] = (mu[i - 1] / mu[i]) ** 2
    return muobs, muexp


def vec(x):
    """vec(R) (src/Jets.jl:66,:758), vec(x) for arrays, and B = vec(A) (:1129-1154): the same operator
    with "vectorised" domain and range, which is what IterativeSolvers-style callers expect.  Device
    vectors are flat already, so vec(A) shares A's operator handle and only relabels the spaces."""
    if isinstance(x, JetAbstractSpace):
        return x.vec()
    if isinstance(x, DeviceArray):
        return x if isinstance(x.space, JetBSpace) else reshape(x, x.space.vec())
    if isinstance(x, JopAdjoint):
        return adjoint(vec(x.op))
    if isinstance(x, Jop):
        if domain(x).ndims == 1 and range_(x).ndims == 1:   # :1130
            return x
        return type(x)(x._h, domain(x).vec(), range_(x).vec(), x.meta)
    raise TypeError(f"vec is not defined for {type(x).__name__}")


def plan_info(A, mode=None):
    e, n = C.c_int32(), C.c_int32()
    check(lib.jets_op_plan_info(A._h.h, A._mode if mode is None else mode, C.byref(e), C.byref(n)))
    names = [nm for bit, nm in ((1, "tma"), (2, "ldg"), (4, "gemv"), (8, "tcgen05"), (16, "staged"), (64, "gather")) if e.value & bit]
    return {"engines": names, "launches": n.value, "input_cache": bool(e.value & 32)}


def launch_count():
    return lib.jets_launch_count()


def set_fused_engine(which):
    check(lib.jets_set_fused_engine({"auto": 0, "tma": 1, "ldg": 2, "tma_nocache": 3}.get(which, which)))
