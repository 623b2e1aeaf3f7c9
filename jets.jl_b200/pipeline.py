"""The host-buffer pipeline (``m_host -> d = A m -> m' = A' d -> m'_host`` over block-row chunks on three
streams) lives in libjets_b200 (csrc/dist_op.cu, ``jets_dist_apply_normal_host``).  This module only exposes
its issue order -- a pure host function of the block structure -- so that the CPU tests can replay it."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import lib

ITEM = {0: "fwd", 1: "adj", 2: "push_prev", 3: "push_next", 4: "partial_prev", 5: "partial_next"}


def schedule(nz, halo, nchunks, has_prev, has_next):
    """``nz``: bool array (nloc, nloc + 2*halo), True where the rank-local operator has a non-zero block.
    Returns (items, chunks, up_need): items = [(name, k)], chunks = [(begin, end)] block rows,
    up_need[k] = last upload chunk the forward of chunk k waits for."""
    nz = np.ascontiguousarray(nz, dtype=np.uint8)
    nloc = nz.shape[0]
    assert nz.shape[1] == nloc + 2 * halo
    cap = 4 * max(nchunks, 1) + 16
    items = (C.c_int32 * (2 * cap))()
    bounds = (C.c_int32 * (2 * cap))()
    need = (C.c_int32 * cap)()
    k = C.c_int32()
    n = lib.jets_dist_pipeline_schedule(nloc, halo, nchunks, int(has_prev), int(has_next), nz.ctypes.data_as(C.c_void_p), cap,
                                        items, bounds, need, C.byref(k))
    if n < 0:
        raise RuntimeError(lib.jets_last_error().decode(errors="replace"))
    return ([(ITEM[items[2 * i]], items[2 * i + 1]) for i in range(n)],
            [(bounds[2 * i], bounds[2 * i + 1]) for i in range(k.value)], [need[i] for i in range(k.value)])
