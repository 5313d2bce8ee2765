"""ctypes declarations for include/burn_b200.h.

This is the Python twin of the Rust `extern "C"` block in INTEGRATION.md: the
same symbols, the same PODs.  It is used by the tests, bench.py and the
host-side mirror; there is no other way into the CUDA library and no fallback
when the library is missing.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

MAX_RANK = 8
NCCL_UNIQUE_ID_BYTES = 128

# b200_dtype
F32, F16, BF16, I32, I64, BOOL, U8 = range(7)
DTYPE_SIZE = {F32: 4, F16: 2, BF16: 2, I32: 4, I64: 8, BOOL: 1, U8: 1}

# b200_status
OK, ERR_CUDA, ERR_INVALID, ERR_SHAPE, ERR_UNSUPPORTED, ERR_NCCL, ERR_NO_DEVICE = 0, -1, -2, -3, -4, -5, -6

# b200_reduce_kind
RED_SUM, RED_MEAN, RED_PROD, RED_MAX, RED_MIN, RED_ARGMAX, RED_ARGMIN, RED_MAXABS, RED_ANY, RED_ALL = range(10)

# b200_mm_precision
MM_TF32, MM_BF16, MM_F32X3 = range(3)

REDUCE_SUM, REDUCE_MEAN = 0, 1

_OPCODES = """MOV ADD_F SUB_F MUL_F DIV_F REM_F POW_F MIN_F MAX_F ATAN2_F NEG_F ABS_F EXP_F LOG_F
LOG1P_F SQRT_F RECIP_F SIN_F COS_F TAN_F TANH_F ERF_F FLOOR_F CEIL_F ROUND_F TRUNC_F SIGN_F SINH_F
COSH_F ASIN_F ACOS_F ATAN_F ASINH_F ACOSH_F ATANH_F SIGMOID_F CLAMP_F EQ_F NE_F LT_F LE_F GT_F GE_F
ISNAN_F ISINF_F ADD_I SUB_I MUL_I DIV_I REM_I MIN_I MAX_I NEG_I ABS_I SIGN_I AND_I OR_I XOR_I NOT_I
SHL_I SHR_I CLAMP_I EQ_I NE_I LT_I LE_I GT_I GE_I AND_B OR_B XOR_B NOT_B SELECT F2I I2F B2F B2I F2B
I2B REMT_F""".split()
OP = {name: i for i, name in enumerate(_OPCODES)}
OP_COUNT = len(_OPCODES)

ARG_ACC = 0x00
DST_NONE = 0xFF


def ARG_INPUT(i: int) -> int:
    return 0x40 | i


def ARG_TEMP(i: int) -> int:
    return 0x80 | i


def ARG_SCALAR(i: int) -> int:
    return 0xC0 | i


class Tensor(C.Structure):
    """b200_tensor"""
    _fields_ = [
        ("ptr", C.c_void_p),
        ("dtype", C.c_int32),
        ("rank", C.c_int32),
        ("shape", C.c_int64 * MAX_RANK),
        ("strides", C.c_int64 * MAX_RANK),
    ]


class TapeOp(C.Structure):
    """b200_tape_op"""
    _fields_ = [
        ("op", C.c_uint8), ("a", C.c_uint8), ("b", C.c_uint8), ("c", C.c_uint8),
        ("dst_temp", C.c_uint8), ("dst_out", C.c_uint8), ("pad", C.c_uint8 * 2),
    ]


class Tape(C.Structure):
    """b200_tape"""
    _fields_ = [
        ("ops", C.POINTER(TapeOp)), ("n_ops", C.c_int32),
        ("scalars", C.POINTER(C.c_uint32)), ("n_scalars", C.c_int32),
    ]


_i32, _i64, _u64, _vp = C.c_int32, C.c_int64, C.c_uint64, C.c_void_p
_TP, _TAPEP = C.POINTER(Tensor), C.POINTER(Tape)

# name -> (restype, argtypes).  Every symbol include/burn_b200.h declares.
SIGNATURES = {
    "b200_abi_version": (_i32, []),
    "b200_device_count": (_i32, [C.POINTER(_i32)]),
    "b200_init": (_i32, [_i32]),
    "b200_set_device": (_i32, [_i32]),
    "b200_device_info": (_i32, [_i32, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_u64)]),
    "b200_stream_create": (_i32, [C.POINTER(_vp), _i32]),
    "b200_stream_destroy": (_i32, [_vp]),
    "b200_stream_sync": (_i32, [_vp]),
    "b200_device_sync": (_i32, []),
    "b200_last_error": (C.c_char_p, []),
    "b200_alloc": (_i32, [C.POINTER(_vp), _u64, _vp]),
    "b200_retain": (_i32, [_vp]),
    "b200_refcount": (_i32, [_vp, C.POINTER(C.c_uint32)]),
    "b200_free": (_i32, [_vp, _vp]),
    "b200_memory_cleanup": (_i32, []),
    "b200_memset": (_i32, [_vp, _i32, _u64, _vp]),
    "b200_host_alloc": (_i32, [C.POINTER(_vp), _u64]),
    "b200_host_free": (_i32, [_vp]),
    "b200_memcpy_h2d": (_i32, [_vp, _vp, _u64, _vp]),
    "b200_memcpy_d2h": (_i32, [_vp, _vp, _u64, _vp]),
    "b200_memcpy_d2d": (_i32, [_vp, _vp, _u64, _vp]),
    "b200_launch_elemwise": (_i32, [_TAPEP, _TP, _i32, _TP, _i32, _i32, C.POINTER(_i64), _vp]),
    "b200_launch_reduce": (_i32, [_i32, _i32, _i32, C.POINTER(_i64), _TAPEP, _TP, _i32, _TAPEP, _TP, _i32, _TP, _i32, _vp]),
    "b200_launch_reduce_full": (_i32, [_i32, _TP, _TP, _vp]),
    "b200_matmul_workspace_bytes": (_i32, [_TP, _TP, _i32, C.POINTER(_u64)]),
    "b200_launch_matmul": (_i32, [_TP, _TP, _TP, _i32, _TAPEP, _TP, _i32, _vp, _u64, _vp]),
    "b200_launch_copy": (_i32, [_TP, _TP, _vp]),
    "b200_launch_gather": (_i32, [_i32, _TP, _TP, _TP, _vp]),
    "b200_launch_scatter_add": (_i32, [_i32, _TP, _TP, _TP, _vp]),
    "b200_launch_select": (_i32, [_i32, _TP, _TP, _TP, _vp]),
    "b200_launch_select_add": (_i32, [_i32, _TP, _TP, _TP, _vp]),
    "b200_launch_slice_assign": (_i32, [_TP, C.POINTER(_i64), C.POINTER(_i64), _TP, _vp]),
    "b200_launch_cat": (_i32, [_TP, _i32, _i32, _TP, _vp]),
    "b200_launch_repeat_dim": (_i32, [_TP, _i32, _i64, _TP, _vp]),
    "b200_launch_flip": (_i32, [_TP, C.POINTER(_i32), _i32, _TP, _vp]),
    "b200_launch_random": (_i32, [_TP, _i32, C.c_double, C.c_double, _u64, _u64, _vp]),
    "b200_launch_arange": (_i32, [_TP, _i64, _i64, _vp]),
    "b200_launch_softmax": (_i32, [_TP, _TP, _i32, _vp]),
    "b200_launch_layer_norm": (_i32, [_TP, _TP, _TP, C.c_double, _TP, _vp]),
    "b200_launch_attention": (_i32, [_TP, _TP, _TP, _TP, C.c_double, C.c_double, _i32, _TP, _TP, _vp]),
    "b200_launch_attention_backward": (_i32, [_TP, _TP, _TP, _TP, _TP, C.c_double, _i32, _TP, _TP, _vp]),
    "b200_launch_attention_flash": (_i32, [_TP, _TP, _TP, _TP, C.c_double, C.c_double, _i32, _TP, _TP, _vp]),
    "b200_launch_attention_flash_backward": (_i32, [_TP, _TP, _TP, _TP, _TP, _TP, _TP, C.c_double, C.c_double, _i32,
                                                    _TP, _TP, _TP, _vp]),
    "b200_launch_softmax_cross_entropy": (_i32, [_TP, _TP, C.c_double, _TP, _TP, _vp]),
    "b200_collective_mark": (_i32, [_vp, _vp]),
    "b200_stream_wait_event": (_i32, [_vp, _vp]),
    "b200_jit_selftest": (_i32, [C.POINTER(C.c_uint64)]),
    "b200_jit_cache_stats": (_i32, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "b200_launch_adam": (_i32, [_TP, _TP, _TP, _TP, _TP, C.c_double, C.c_double, C.c_double, _vp]),
    "b200_launch_softmax_backward": (_i32, [_TP, _TP, _TP, C.c_double, _TP, _vp]),
    "b200_layer_norm_backward_partials": (_i32, [_TP, C.POINTER(C.c_int32)]),
    "b200_launch_layer_norm_backward": (_i32, [_TP, _TP, _TP, C.c_double, _TP, _TP, _TP, _vp]),
    "b200_launch_layer_norm_backward_ex": (_i32, [_TP, _TP, _TP, C.c_double, _TP, _TP, _TP, _TP, _vp]),
    "b200_comm_unique_id": (_i32, [C.POINTER(C.c_uint8)]),
    "b200_comm_init": (_i32, [C.POINTER(_vp), C.POINTER(C.c_uint8), _i32, _i32]),
    "b200_comm_destroy": (_i32, [_vp]),
    "b200_all_reduce": (_i32, [_vp, _vp, _u64, _i32, _i32, _vp]),
    "b200_all_reduce_multi": (_i32, [_vp, C.POINTER(_vp), C.POINTER(_u64), _i32, _i32, _i32, _vp]),
    "b200_collective_sync": (_i32, [_vp, _vp]),
    "b200_comm_init_all": (_i32, [C.POINTER(_vp), C.POINTER(_i32), _i32]),
    "b200_all_reduce_group": (_i32, [C.POINTER(_vp), C.POINTER(_vp), _u64, _i32, _i32, _i32, C.POINTER(_vp)]),
    "b200_comm_host_sync": (_i32, [_vp]),
    "b200_peer_alloc": (_i32, [C.POINTER(_vp), _u64]),
    "b200_peer_free": (_i32, [_vp]),
    "b200_peer_export": (_i32, [_vp, C.POINTER(C.c_uint8)]),
    "b200_peer_flag_bytes": (_u64, []),
    "b200_peer_group_create": (_i32, [C.POINTER(_vp), _i32, _i32, _vp, _u64, C.POINTER(C.c_uint8)]),
    "b200_peer_group_create_local": (_i32, [C.POINTER(_vp), C.POINTER(_i32), _i32, _u64]),
    "b200_peer_data": (_vp, [_vp]),
    "b200_peer_group_destroy": (_i32, [_vp]),
    "b200_launch_peer_all_reduce": (_i32, [_vp, _u64, _u64, _i32, _i32, _vp]),
    "b200_launch_peer_adam": (_i32, [_vp, _u64, _u64, _vp, _vp, _vp, _u64, C.c_double, C.c_double, C.c_double, _i32, _vp]),
    "b200_peer_sync": (_i32, [_vp, _vp]),
    "b200_peer_mark": (_i32, [_vp, _vp]),
    "b200_peer_host_sync": (_i32, [_vp]),
    "b200_event_create": (_i32, [C.POINTER(_vp)]),
    "b200_event_destroy": (_i32, [_vp]),
    "b200_event_record": (_i32, [_vp, _vp]),
    "b200_event_query": (_i32, [_vp, C.POINTER(_i32)]),
    "b200_event_elapsed_ms": (_i32, [_vp, _vp, C.POINTER(C.c_float)]),
    "b200_graph_begin": (_i32, [_vp]),
    "b200_graph_end": (_i32, [_vp, C.POINTER(C.c_void_p)]),
    "b200_graph_launch": (_i32, [_vp, _vp]),
    "b200_graph_destroy": (_i32, [_vp]),
    "b200_graph_node_count": (_i32, [_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "b200_launch_count": (_u64, []),
    "b200_launch_count_reset": (None, []),
}

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libburn_b200.so"


class B200Error(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"b200 status {status}: {message}")
        self.status = status
        self.message = message


_lib = None


def load() -> C.CDLL:
    """Loads libburn_b200.so and binds every declared symbol.  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("BURN_B200_LIB", LIB_PATH))
    if not path.exists():
        raise ImportError(
            f"{path} is missing — build it with `python -m burn_b200.build` "
            "(burn-b200 has no CPU or PyTorch fallback)")
    lib = C.CDLL(str(path), mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        raise B200Error(status, load().b200_last_error().decode("utf-8", "replace"))
