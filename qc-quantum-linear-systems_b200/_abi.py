"""ctypes binding of ``include/qls_abi.h`` (libqls_b200.so).

There is no fallback: if the library is missing or a call fails, an exception is raised."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

QLS_C128, QLS_C64 = 0, 1
QLS_OP_DENSE, QLS_OP_DIAG, QLS_OP_MUXRY, QLS_OP_MATVEC = 1, 2, 3, 4
PARAM_KINDS = {"ry": 0, "rx": 1, "rz": 2, "p": 3}
QLS_MAX_SEG, QLS_MAX_HIGH, QLS_MAX_FSEG = 6, 8, 16
QLS_OP_FLAG_REAL, QLS_OP_FLAG_PARAM = 1, 2
QLS_OP_FLAG_ATTACH_PRE, QLS_OP_FLAG_ATTACH_POST, QLS_MAX_ATTACH = 4, 8, 4
QLS_RED_NORM_RANGE, QLS_RED_PROB_MASK, QLS_RED_EXPECT_Z, QLS_RED_INNER = 0, 1, 2, 3
QLS_GEMM_CHUNK_ROWS = 1 << 14  # include/qls_abi.h
QLS_REDUCE_WORKSPACE_BYTES = 64 * 1024
UINT64_MAX = 2**64 - 1

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libqls_b200.so")


class QlsSeg(C.Structure):
    _fields_ = [("src", C.c_uint8), ("width", C.c_uint8), ("dst", C.c_uint8), ("pad", C.c_uint8)]


class QlsOp(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("k", C.c_int32),
        ("tq", C.c_uint8 * 8),
        ("lc_mask", C.c_uint32),
        ("lc_val", C.c_uint32),
        ("xc_mask", C.c_uint64),
        ("xc_val", C.c_uint64),
        ("pool_off", C.c_uint64),
        ("n_lseg", C.c_uint8),
        ("n_xseg", C.c_uint8),
        ("flags", C.c_uint8),
        ("pad0", C.c_uint8),
        ("pad1", C.c_uint32),
        ("lseg", QlsSeg * QLS_MAX_SEG),
        ("xseg", QlsSeg * QLS_MAX_SEG),
        ("param_index", C.c_uint32),
        ("param_coeff", C.c_float),
        ("param_offset", C.c_float),
        ("reserved", C.c_uint8 * 12),
    ]


class QlsPass(C.Structure):
    _fields_ = [
        ("tile_bits", C.c_int32),
        ("low_bits", C.c_int32),
        ("n_high", C.c_int32),
        ("high", C.c_uint8 * QLS_MAX_HIGH),
        ("op_begin", C.c_uint32),
        ("op_count", C.c_uint32),
        ("fix_mask", C.c_uint64),
        ("fix_val", C.c_uint64),
        ("n_fseg", C.c_int32),
        ("fseg", QlsSeg * QLS_MAX_FSEG),
        ("pad", C.c_uint32),
        ("rs_begin", C.c_uint32),
        ("rs_count", C.c_uint32),
    ]


QLS_RS_STAGE, QLS_RS_G1, QLS_RS_G2, QLS_RS_DIAG, QLS_RS_MUXRY = 1, 2, 3, 4, 5


class QlsRsOp(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("r", C.c_uint8 * 4),
        ("ctl_tmask", C.c_uint32),
        ("ctl_tval", C.c_uint32),
        ("ctl_rmask", C.c_uint32),
        ("ctl_rval", C.c_uint32),
        ("xc_mask", C.c_uint64),
        ("xc_val", C.c_uint64),
        ("pool_off", C.c_uint64),
        ("n_tseg", C.c_uint8),
        ("n_xseg", C.c_uint8),
        ("flags", C.c_uint8),
        ("pad0", C.c_uint8),
        ("pad1", C.c_uint32),
        ("tseg", QlsSeg * QLS_MAX_SEG),
        ("xseg", QlsSeg * QLS_MAX_SEG),
        ("rt", C.c_uint16 * 32),
        ("reserved", C.c_uint8 * 24),
    ]


assert C.sizeof(QlsOp) == 128, C.sizeof(QlsOp)
assert C.sizeof(QlsPass) == 128, C.sizeof(QlsPass)
assert C.sizeof(QlsRsOp) == 192, C.sizeof(QlsRsOp)

# every symbol declared in include/qls_abi.h: name -> (restype, argtypes)
_vp, _u64, _u32, _i = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
SYMBOLS = {
    "qls_abi_version": (_i, []),
    "qls_last_error": (C.c_char_p, []),
    "qls_device_info": (_i, [_i, C.POINTER(_i), C.POINTER(_u64), C.POINTER(_u64), C.POINTER(_u64)]),
    "qls_program_create": (_i, [C.POINTER(QlsPass), _u32, C.POINTER(QlsOp), _u32, C.POINTER(QlsRsOp), _u32, _vp, _u64, _i, _i,
                                C.POINTER(_vp)]),
    "qls_set_sweep_only": (_i, [_i]),
    "qls_program_destroy": (_i, [_vp]),
    "qls_jit_compile": (_i, [C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), _i, C.POINTER(C.c_char_p), _i,
                            C.POINTER(_vp), C.POINTER(_u64), C.c_char_p, _u64]),
    "qls_jit_free": (_i, [_vp]),
    "qls_program_set_pass_kernel": (_i, [_vp, _u32, _vp, _u64, C.c_char_p, _u32, _u32]),
    "qls_program_set_pass_tma": (_i, [_vp, _u32, _i, _i, C.POINTER(C.c_int), C.POINTER(C.c_int), _i]),
    "qls_set_use_jit": (_i, [_i]),
    "qls_set_trace_buffer": (_i, [_vp]),
    "qls_program_run": (_i, [_vp, _vp, _i, _u64, _u32, _u32, _vp]),
    "qls_sv_init_basis": (_i, [_vp, _i, _i, _u64, _vp]),
    "qls_sv_write": (_i, [_vp, _i, _u64, _vp, _u64, _vp]),
    "qls_sv_read": (_i, [_vp, _i, _u64, _vp, _u64, _vp]),
    "qls_sv_scale": (_i, [_vp, _i, _i, C.c_double, C.c_double, _vp]),
    "qls_reduce": (_i, [_vp, _i, _i, _i, _u64, _u64, _u64, _vp, _vp, C.POINTER(C.c_double), _vp]),
    "qls_ctrl_gemm": (_i, [_vp, _vp, _i, _i, _i, _u64, _u64, _u64, _vp, _vp]),
    "qls_batch_expect_z": (_i, [_vp, _u32, _i, _i, _vp, _u32, _u32, _vp, _vp]),
    "qls_pack_bit": (_i, [_vp, _vp, _i, _i, _i, _i, _u64, _u64, _vp]),
    "qls_unpack_bit": (_i, [_vp, _vp, _i, _i, _i, _i, _u64, _u64, _vp]),
    "qls_pack_bits": (_i, [_vp, _vp, _i, _i, C.POINTER(C.c_int), _i, _u64, _u64, _u64, _vp]),
    "qls_unpack_bits": (_i, [_vp, _vp, _i, _i, C.POINTER(C.c_int), _i, _u64, _u64, _u64, _vp]),
    "qls_enable_peer_access": (_i, [_i, _i]),
    "qls_ipc_open": (_i, [_vp, _i, C.POINTER(C.c_void_p)]),
    "qls_ipc_close": (_i, [_vp]),
    "qls_swap_peer": (_i, [_vp, _vp, _i, _i, C.POINTER(C.c_int), _i, _u64, _u64, _u64, _u64, _vp]),
    "qls_copy_async": (_i, [_vp, _vp, _u64, _vp]),
    "qls_copy2d_async": (_i, [_vp, _u64, _vp, _u64, _u64, _u64, _vp]),
    "qls_program_run_part": (_i, [_vp, _vp, _i, _u64, C.c_uint32, _i, _i, _i, _i, _vp]),
    "qls_set_pass_grid_reserve": (_i, [_i]),
}

_lib: Optional[C.CDLL] = None


class QlsError(RuntimeError):
    pass


def lib_path() -> str:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load libqls_b200.so and bind every declared symbol.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise QlsError(
            f"CUDA library not built: {_LIB_PATH} is missing. Run `python __graft_entry__.py build` "
            "(nvcc, sm_100a). There is no CPU fallback."
        )
    lib = C.CDLL(_LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.qls_abi_version() != 1:
        raise QlsError("ABI version mismatch between _abi.py and libqls_b200.so")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().qls_last_error()
        raise QlsError(f"libqls_b200 error {rc}: {msg.decode() if msg else '?'}")
