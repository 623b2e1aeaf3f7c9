"""ctypes access to oracle/libjets_oracle.so (C restatement of the Jets CPU hot loops).
TEST / MEASUREMENT INFRASTRUCTURE ONLY: imported by tests/ and by bench.py's CPU-baseline legs."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libjets_oracle.so")

LEAF = {"zero": 0, "diag": 1, "fdiff": 2, "lap": 3}


class Leaf(C.Structure):
    _fields_ = [("kind", C.c_int32), ("pad", C.c_int32), ("state", C.c_void_p)]


def load():
    if not os.path.exists(SO):
        subprocess.run(["make", "-s", "-C", HERE], check=True)
    lib = C.CDLL(SO)
    for suf in ("f32", "f64"):
        f = getattr(lib, f"jets_ref_block_apply_{suf}")
        f.restype = C.c_int
        f.argtypes = [C.c_int32, C.c_int32, C.POINTER(Leaf), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                      C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        f = getattr(lib, f"jets_ref_chain_apply_{suf}")
        f.restype = C.c_int
        f.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        f = getattr(lib, f"jets_ref_dot_{suf}")
        f.restype = C.c_double
        f.argtypes = [C.c_int32, C.POINTER(C.c_int64), C.c_void_p, C.c_void_p, C.c_int]
    lib.jets_ref_num_threads.restype = C.c_int
    return lib


def _suf(dt):
    return "f32" if np.dtype(dt) == np.float32 else "f64"


class BlockOp:
    """R x C block operator of leaves; `leaves[r][c]` is ("zero",None) | ("diag", w) | ("fdiff",None)
    | ("lap",None)."""

    def __init__(self, leaves, dom_len, rng_len, dtype):
        self.lib = load()
        self.R, self.Cn = len(leaves), len(leaves[0])
        self.dtype = np.dtype(dtype)
        self.keep = []
        arr = (Leaf * (self.R * self.Cn))()
        for c in range(self.Cn):
            for r in range(self.R):
                kind, st = leaves[r][c]
                e = arr[r + c * self.R]
                e.kind = LEAF[kind]
                if st is not None:
                    st = np.ascontiguousarray(st, dtype=self.dtype)
                    self.keep.append(st)
                    e.state = st.ctypes.data
        self.arr = arr
        self.dom_len = (C.c_int64 * self.Cn)(*dom_len)
        self.rng_len = (C.c_int64 * self.R)(*rng_len)
        self.ndom, self.nrng = int(sum(dom_len)), int(sum(rng_len))

    def apply(self, x, adj=False, mode=0, out=None):
        x = np.ascontiguousarray(x, dtype=self.dtype)
        out = np.empty(self.ndom if adj else self.nrng, dtype=self.dtype) if out is None else out
        f = getattr(self.lib, f"jets_ref_block_apply_{_suf(self.dtype)}")
        f(self.R, self.Cn, self.arr, self.dom_len, self.rng_len, x.ctypes.data, out.ctypes.data,
          1 if adj else 0, mode)
        return out


def chain_apply(w, mo, x, adj=False, mode=0, out=None):
    lib = load()
    out = np.empty_like(x) if out is None else out
    getattr(lib, f"jets_ref_chain_apply_{_suf(x.dtype)}")(x.size, w.ctypes.data, mo.ctypes.data, x.ctypes.data,
                                                           out.ctypes.data, 1 if adj else 0, mode)
    return out


def num_threads():
    return load().jets_ref_num_threads()


def use_all_cores():
    """Sets the OpenMP team to every core this process may run on (ignores OMP_NUM_THREADS) and returns it."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib = load()
    lib.jets_ref_set_threads(int(n))
    return lib.jets_ref_num_threads()
