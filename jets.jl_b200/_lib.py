"""ctypes binding of libjets_b200.so (include/jets_b200.h).

The library is the product: if it is missing or cannot be loaded this module raises -- there is
no Python/numpy/torch fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("JETS_B200_LIB") or os.path.join(HERE, "libjets_b200.so")   # override: A/B against another build

F32, F64, C64, C128 = 0, 1, 2, 3
MODE_F, MODE_DF, MODE_DFT = 0, 1, 2
PW = {"square": 0, "power": 1, "exp": 2, "sin": 3, "tanh": 4, "log": 5, "atan": 6}
STENCIL = {"fdiff": 0, "lap": 1}
COEF_NEG, COEF_INV = 1, 2

STATUS = {0: "OK", 1: "INVALID", 2: "SHAPE", 3: "DTYPE", 4: "CUDA", 5: "UNSUPPORTED", 6: "NOT_LINEAR",
          7: "NO_POINT", 8: "NCCL"}


class JetsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"jets_b200[{STATUS.get(code, code)}]: {msg}")
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(jets_b200 has no CPU fallback)")
    return C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)


lib = _load()

_p, _i, _i32, _i64, _u64, _d = C.c_void_p, C.c_int, C.c_int32, C.c_int64, C.c_uint64, C.c_double
_pp = C.POINTER(C.c_void_p)
_pi64 = C.POINTER(C.c_int64)
_pd = C.POINTER(C.c_double)
_pi32 = C.POINTER(C.c_int32)

# name -> (restype, argtypes); every symbol declared in include/jets_b200.h appears here
SIGNATURES = {
    "jets_abi_version": (_i, []),
    "jets_init": (_i, [_i]),
    "jets_shutdown": (_i, []),
    "jets_last_error": (C.c_char_p, []),
    "jets_stream_set": (_i, [_p]),
    "jets_stream_fork": (_i, [_i]),
    "jets_stream_main": (_i, []),
    "jets_stream_join": (_i, [_i]),
    "jets_stream_get": (_p, []),
    "jets_sync": (_i, []),
    "jets_launch_count": (_i64, []),
    "jets_debug_trace": (_i64, [_p, _i64]),
    "jets_device_sm_count": (_i, []),
    "jets_buf_create": (_i, [_i, _i32, _pi64, _pp]),
    "jets_buf_wrap": (_i, [_i, _p, _i32, _pi64, _pp]),
    "jets_buf_view": (_i, [_p, _i32, _i32, _pp]),
    "jets_buf_reshape": (_i, [_p, _i32, _pi64, _pp]),
    "jets_buf_retain": (_i, [_p]),
    "jets_buf_destroy": (_i, [_p]),
    "jets_buf_dtype": (_i, [_p]),
    "jets_buf_nblocks": (_i32, [_p]),
    "jets_buf_length": (_i64, [_p]),
    "jets_buf_block_range": (_i, [_p, _i32, _pi64, _pi64]),
    "jets_buf_devptr": (_p, [_p]),
    "jets_buf_upload": (_i, [_p, _i32, _p, _i64]),
    "jets_buf_download": (_i, [_p, _i32, _p, _i64]),
    "jets_buf_upload_async": (_i, [_p, _i32, _p, _i64]),
    "jets_buf_download_async": (_i, [_p, _i32, _p, _i64]),
    "jets_buf_write": (_i, [_p, _i64, _p, _i64]),
    "jets_buf_read": (_i, [_p, _i64, _p, _i64]),
    "jets_buf_copy": (_i, [_p, _p]),
    "jets_buf_fill_c": (_i, [_p, _d, _d]),
    "jets_dot_c": (_i, [_p, _p, _pd]),
    "jets_norm_weighted": (_i, [_p, _p, _d, _pd]),
    "jets_abs": (_i, [_p, _p]),
    "jets_lincomb_c": (_i, [_p, _i32, _pd, _pp]),
    "jets_op_scale_c": (_i, [_i, _i64, _d, _d, _pp]),
    "jets_op_scalar_mul_c": (_i, [_d, _d, _p, _pp]),
    "jets_buf_fill": (_i, [_p, _d]),
    "jets_buf_rand": (_i, [_p, _u64, _u64, _i]),
    "jets_dot": (_i, [_p, _p, _pd]),
    "jets_norm": (_i, [_p, _d, _pd]),
    "jets_extrema": (_i, [_p, _pd, _pd]),
    "jets_lincomb": (_i, [_p, _i32, _pd, _pp]),
    "jets_hadamard": (_i, [_p, _p, _p]),
    "jets_scalar_create": (_i, [_pp]),
    "jets_scalar_destroy": (_i, [_p]),
    "jets_scalar_set": (_i, [_p, _d]),
    "jets_scalar_get": (_i, [_p, _pd]),
    "jets_dot_dev": (_i, [_p, _p, _p]),
    "jets_norm_dev": (_i, [_p, _d, _p]),
    "jets_scalar_op": (_i, [_p, C.c_char, _p, _p]),
    "jets_scalar_prog": (_i, [_i32, _p, C.c_char_p, _p, _p]),
    "jets_apply_axpby": (_i, [_p, _i, _p, _p, _p, _d, _i, _p, _d, _i]),
    "jets_axpby_dev": (_i, [_p, _p, _d, _i, _p, _p, _d, _i, _p]),
    "jets_axpby_pair_dev": (_i, [_p, _p, _d, _i, _p, _p, _d, _i, _p, _p, _p, _d, _i, _p, _p, _d, _i, _p]),
    "jets_apply_axpby_norm": (_i, [_p, _i, _p, _p, _p, _d, _i, _p, _d, _i, _p]),
    "jets_graph_begin": (_i, []),
    "jets_graph_end": (_i, [_pp]),
    "jets_graph_launch": (_i, [_p]),
    "jets_graph_destroy": (_i, [_p]),
    "jets_op_diag": (_i, [_p, _pp]),
    "jets_op_scale": (_i, [_i, _i64, _d, _pp]),
    "jets_op_pointwise": (_i, [_i, _i64, _i, _d, _pp]),
    "jets_op_stencil": (_i, [_i, _i64, _i, _pp]),
    "jets_op_dense": (_i, [_p, _i64, _i64, _i64, _pp]),
    "jets_op_zero": (_i, [_i, _i64, _i64, _pp]),
    "jets_op_restrict": (_i, [_i, _i64, _i64, _pi64, _pp]),
    "jets_op_as_linear": (_i, [_p, _pp]),
    "jets_op_adjoint": (_i, [_p, _pp]),
    "jets_op_compose": (_i, [_i32, _pp, _pp]),
    "jets_op_sum": (_i, [_i32, _pp, _pi32, _pp]),
    "jets_op_block": (_i, [_i32, _i32, _pp, _i, _pp]),
    "jets_op_scalar_mul": (_i, [_d, _p, _pp]),
    "jets_op_retain": (_i, [_p]),
    "jets_op_destroy": (_i, [_p]),
    "jets_op_is_linear": (_i, [_p]),
    "jets_op_is_zero": (_i, [_p]),
    "jets_op_is_block": (_i, [_p]),
    "jets_op_dtype": (_i, [_p]),
    "jets_op_nblocks": (_i32, [_p, _i]),
    "jets_op_block_len": (_i, [_p, _i, _i32, _pi64]),
    "jets_op_getblock": (_i, [_p, _i32, _i32, _pp]),
    "jets_op_set_point": (_i, [_p, _p]),
    "jets_op_jacobian": (_i, [_p, _p, _pp]),
    "jets_op_clone": (_i, [_p, _pp]),
    "jets_apply": (_i, [_p, _i, _p, _p, _i]),
    "jets_op_plan_info": (_i, [_p, _i, _pi32, _pi32]),
    "jets_set_fused_engine": (_i, [_i]),
    "jets_dist_unique_id": (_i, [C.c_char_p]),
    "jets_dist_init": (_i, [_i, _i, C.c_char_p]),
    "jets_dist_init_host": (_i, [_i, _i, _p, _p]),
    "jets_dist_shutdown": (_i, []),
    "jets_dist_rank": (_i, []),
    "jets_dist_size": (_i, []),
    "jets_dist_sum_scalar": (_i, [_pd]),
    "jets_dist_allgather": (_i, [_p, _p]),
    "jets_dist_reduce_scatter": (_i, [_p, _p]),
    "jets_dist_op_create": (_i, [_p, _i32, _pp]),
    "jets_dist_op_create_dense": (_i, [_p, _pp]),
    "jets_dist_op_destroy": (_i, [_p]),
    "jets_dist_apply": (_i, [_p, _i, _p, _p]),
    "jets_dist_op_register": (_i, [_p, _p]),
    "jets_dist_apply_normal_host": (_i, [_p, _p, _p, _i32]),
    "jets_dist_op_join": (_i, [_p]),
    "jets_dist_op_info": (_i32, [_p, _i32]),
    "jets_dist_pipeline_schedule": (_i32, [_i32, _i32, _i32, _i32, _i32, _p, _i32, _pi32, _pi32, _pi32, _pi32]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _f = getattr(lib, _name)  # AttributeError here = the .so does not export a declared symbol
    _f.restype = _res
    _f.argtypes = _args


def check(rc: int):
    if rc != 0:
        raise JetsError(rc, lib.jets_last_error().decode(errors="replace"))


_initialized = {"device": None}


def init(device: int | None = None):
    """jets_init: bind this process to one GPU.  Raises JetsError[CUDA] without a B200."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if _initialized["device"] == device:
        return
    check(lib.jets_init(device))
    _initialized["device"] = device


def ensure_init():
    if _initialized["device"] is None:
        init()
