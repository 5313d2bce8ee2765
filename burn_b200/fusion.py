"""ctypes face of the C++ host fusion layer (include/burn_b200_host.h, burn_b200/host/fusion.cpp).

`FusionStream` plays the role of `Fusion<B>`'s per-device stream: ops are recorded lazily as
OperationIr-like entries, the ElementWise / Matmul / Reduce fusers carve the queue into blocks and
every block runs as one kernel when the stream is drained (`read`, `sync`).  `LazyTensor.drop()`
is `OperationIr::Drop` — dropping an intermediate before the drain is what lets it live in
registers only, exactly as with burn-fusion's refcounted `FusionTensor`
(crates/burn-fusion/src/tensor.rs:108-156).
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _abi as abi
from ._abi import check

BLOCK_ELEMWISE, BLOCK_REDUCE, BLOCK_MATMUL, BLOCK_EAGER, BLOCK_ROWNORM, BLOCK_VIEW = range(6)
FUSER_OPEN, FUSER_CLOSED = 0, 1


class BlockInfo(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_ops", C.c_int32), ("n_inputs", C.c_int32), ("n_outputs", C.c_int32),
                ("n_tape_ops", C.c_int32), ("launches", C.c_int32), ("aliased", C.c_int32), ("from_cache", C.c_int32),
                ("score", C.c_uint64)]


class CacheStats(C.Structure):
    _fields_ = [("hits", C.c_uint64), ("misses", C.c_uint64), ("plans", C.c_uint64), ("inplace_aliases", C.c_uint64)]


_vp, _i32, _i64 = C.c_void_p, C.c_int32, C.c_int64
HOST_SIGNATURES = {
    "b200h_stream_create": (_i32, [C.POINTER(_vp), _i32]),
    "b200h_stream_destroy": (_i32, [_vp]),
    "b200h_from_host": (_i64, [_vp, _vp, _i32, _i32, C.POINTER(_i64)]),
    "b200h_read": (_i32, [_vp, _i64, _vp, C.c_uint64]),
    "b200h_shape": (_i32, [_vp, _i64, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_i64)]),
    "b200h_sync": (_i32, [_vp]),
    "b200h_flush": (_i32, [_vp]),
    "b200h_reshape": (_i64, [_vp, _i64, _i32, C.POINTER(_i64)]),
    "b200h_expand": (_i64, [_vp, _i64, _i32, C.POINTER(_i64)]),
    "b200h_slice": (_i64, [_vp, _i64, C.POINTER(_i64), C.POINTER(_i64)]),
    "b200h_gather": (_i64, [_vp, _i32, _i64, _i64]),
    "b200h_select": (_i64, [_vp, _i32, _i64, _i64]),
    "b200h_from_device": (_i64, [_vp, _vp, _i32, _i32, C.POINTER(_i64), C.POINTER(_i64)]),
    "b200h_device_tensor": (_i32, [_vp, _i64, C.POINTER(abi.Tensor)]),
    "b200h_cache_stats_get": (_i32, [_vp, C.POINTER(CacheStats)]),
    "b200h_cache_clear": (_i32, [_vp]),
    "b200h_fuser_create": (_i32, [_vp, _i32, C.POINTER(_vp)]),
    "b200h_fuser_destroy": (_i32, [_vp]),
    "b200h_fuser_fuse_next": (_i32, [_vp]),
    "b200h_fuser_status_get": (_i32, [_vp]),
    "b200h_fuser_properties": (_i32, [_vp, C.POINTER(C.c_uint64), C.POINTER(_i32)]),
    "b200h_fuser_len": (_i32, [_vp]),
    "b200h_fuser_reset": (_i32, [_vp]),
    "b200h_fuser_clone": (_i32, [_vp, C.POINTER(_vp)]),
    "b200h_fuser_finish": (_i32, [_vp, C.POINTER(_vp)]),
    "b200h_optimization_destroy": (_i32, [_vp]),
    "b200h_optimization_len": (_i32, [_vp]),
    "b200h_optimization_name": (C.c_char_p, [_vp]),
    "b200h_optimization_execute": (_i32, [_vp, _vp]),
    "b200h_optimization_to_state": (_i32, [_vp, _vp, C.c_uint64, C.POINTER(C.c_uint64)]),
    "b200h_optimization_from_state": (_i32, [_vp, C.c_uint64, C.POINTER(_vp)]),
    "b200h_binary": (_i64, [_vp, _i32, _i64, _i64]),
    "b200h_scalar": (_i64, [_vp, _i32, _i64, C.c_double]),
    "b200h_unary": (_i64, [_vp, _i32, _i64]),
    "b200h_mask_fill": (_i64, [_vp, _i64, _i64, C.c_double]),
    "b200h_mask_where": (_i64, [_vp, _i64, _i64, _i64]),
    "b200h_reduce_dim": (_i64, [_vp, _i32, _i64, _i32]),
    "b200h_matmul": (_i64, [_vp, _i64, _i64, _i32]),
    "b200h_swap_dims": (_i64, [_vp, _i64, _i32, _i32]),
    "b200h_drop": (_i32, [_vp, _i64]),
    "b200h_block_count": (_i32, [_vp]),
    "b200h_block_get": (_i32, [_vp, _i32, C.POINTER(BlockInfo)]),
    "b200h_block_clear": (_i32, [_vp]),
}

_NP = {abi.F32: np.float32, abi.I32: np.int32, abi.I64: np.int64, abi.BOOL: np.uint8, abi.U8: np.uint8,
       abi.F16: np.float16}
_bound = False


def _lib():
    global _bound
    lib = abi.load()
    if not _bound:
        for name, (res, args) in HOST_SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _bound = True
    return lib


def _id(v: int) -> int:
    if v < 0:
        raise abi.B200Error(-1, abi.load().b200_last_error().decode("utf-8", "replace"))
    return v


class LazyTensor:
    def __init__(self, stream: "FusionStream", tid: int):
        self.stream, self.id, self._dropped = stream, tid, False

    # ---- metadata
    def _meta(self):
        dt, rk = C.c_int32(), C.c_int32()
        shape = (C.c_int64 * abi.MAX_RANK)()
        check(_lib().b200h_shape(self.stream.h, self.id, C.byref(dt), C.byref(rk), shape))
        return dt.value, tuple(shape[i] for i in range(rk.value))

    @property
    def shape(self):
        return self._meta()[1]

    @property
    def dtype(self):
        return self._meta()[0]

    # ---- ownership
    def drop(self) -> None:
        if not self._dropped:
            self._dropped = True
            check(_lib().b200h_drop(self.stream.h, self.id))

    # ---- readback (drains the stream)
    def numpy(self) -> np.ndarray:
        dt, shape = self._meta()
        out = np.empty(shape, dtype=_NP[dt])
        check(_lib().b200h_read(self.stream.h, self.id, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out.astype(bool) if dt == abi.BOOL else out

    # ---- ops named after the reference Tensor API
    def _bin(self, op, other): return LazyTensor(self.stream, _id(_lib().b200h_binary(self.stream.h, abi.OP[op], self.id, other.id)))
    def _sc(self, op, s): return LazyTensor(self.stream, _id(_lib().b200h_scalar(self.stream.h, abi.OP[op], self.id, float(s))))
    def _un(self, op): return LazyTensor(self.stream, _id(_lib().b200h_unary(self.stream.h, abi.OP[op], self.id)))

    def add(self, o): return self._bin("ADD_F", o)
    def sub(self, o): return self._bin("SUB_F", o)
    def mul(self, o): return self._bin("MUL_F", o)
    def div(self, o): return self._bin("DIV_F", o)
    def add_scalar(self, s): return self._sc("ADD_F", s)
    def sub_scalar(self, s): return self._sc("SUB_F", s)
    def mul_scalar(self, s): return self._sc("MUL_F", s)
    def div_scalar(self, s): return self._sc("DIV_F", s)
    def lower_equal_elem(self, s): return self._sc("LE_F", s)
    def lower_elem(self, s): return self._sc("LT_F", s)
    def greater_elem(self, s): return self._sc("GT_F", s)
    def exp(self): return self._un("EXP_F")
    def log(self): return self._un("LOG_F")
    def erf(self): return self._un("ERF_F")
    def sqrt(self): return self._un("SQRT_F")
    def tanh(self): return self._un("TANH_F")
    def mask_fill(self, mask, value): return LazyTensor(self.stream, _id(_lib().b200h_mask_fill(self.stream.h, self.id, mask.id, float(value))))
    def mask_where(self, mask, src): return LazyTensor(self.stream, _id(_lib().b200h_mask_where(self.stream.h, self.id, mask.id, src.id)))
    def _red(self, kind, dim): return LazyTensor(self.stream, _id(_lib().b200h_reduce_dim(self.stream.h, kind, self.id, dim)))
    def sum_dim(self, dim): return self._red(abi.RED_SUM, dim)
    def mean_dim(self, dim): return self._red(abi.RED_MEAN, dim)
    def max_dim(self, dim): return self._red(abi.RED_MAX, dim)
    def argmax(self, dim): return self._red(abi.RED_ARGMAX, dim)
    def matmul(self, o, precision=abi.MM_F32X3): return LazyTensor(self.stream, _id(_lib().b200h_matmul(self.stream.h, self.id, o.id, precision)))
    def swap_dims(self, d0, d1): return LazyTensor(self.stream, _id(_lib().b200h_swap_dims(self.stream.h, self.id, d0, d1)))

    def reshape(self, shape):
        sh = (C.c_int64 * len(shape))(*shape)
        return LazyTensor(self.stream, _id(_lib().b200h_reshape(self.stream.h, self.id, len(shape), sh)))

    def expand(self, shape):
        sh = (C.c_int64 * len(shape))(*shape)
        return LazyTensor(self.stream, _id(_lib().b200h_expand(self.stream.h, self.id, len(shape), sh)))

    def slice(self, ranges):
        """ranges: one (start, end) per dim, already canonical."""
        st = (C.c_int64 * len(ranges))(*[r[0] for r in ranges])
        en = (C.c_int64 * len(ranges))(*[r[1] for r in ranges])
        return LazyTensor(self.stream, _id(_lib().b200h_slice(self.stream.h, self.id, st, en)))

    def gather(self, dim, indices): return LazyTensor(self.stream, _id(_lib().b200h_gather(self.stream.h, dim, self.id, indices.id)))
    def select(self, dim, indices): return LazyTensor(self.stream, _id(_lib().b200h_select(self.stream.h, dim, self.id, indices.id)))

    def device_tensor(self) -> "abi.Tensor":
        """Drain the stream and return the kernel-ABI descriptor of this tensor's storage."""
        d = abi.Tensor()
        check(_lib().b200h_device_tensor(self.stream.h, self.id, C.byref(d)))
        return d


class FusionStream:
    def __init__(self, plan_only: bool = False):
        self.h = C.c_void_p()
        check(_lib().b200h_stream_create(C.byref(self.h), 1 if plan_only else 0))
        self.plan_only = plan_only

    def close(self):
        if self.h:
            _lib().b200h_stream_destroy(self.h)
            self.h = C.c_void_p()

    def tensor(self, a) -> LazyTensor:
        a = np.ascontiguousarray(a)
        dt = {np.dtype(np.float32): abi.F32, np.dtype(np.int32): abi.I32, np.dtype(np.int64): abi.I64,
              np.dtype(np.bool_): abi.BOOL, np.dtype(np.uint8): abi.U8}[a.dtype]
        if a.dtype == np.bool_:
            a = a.astype(np.uint8)
        shape = (C.c_int64 * max(a.ndim, 1))(*a.shape)
        return LazyTensor(self, _id(_lib().b200h_from_host(self.h, a.ctypes.data_as(C.c_void_p), dt, a.ndim, shape)))

    def placeholder(self, shape: Sequence[int], dtype=abi.F32) -> LazyTensor:
        """A tensor with no host data (plan-only streams)."""
        sh = (C.c_int64 * len(shape))(*shape)
        return LazyTensor(self, _id(_lib().b200h_from_host(self.h, None, dtype, len(shape), sh)))

    def wrap(self, desc: "abi.Tensor") -> LazyTensor:
        """Non-owning handle on device memory the caller owns (a DeviceTensor's descriptor)."""
        sh = (C.c_int64 * desc.rank)(*desc.shape[:desc.rank])
        st = (C.c_int64 * desc.rank)(*desc.strides[:desc.rank])
        return LazyTensor(self, _id(_lib().b200h_from_device(self.h, desc.ptr, desc.dtype, desc.rank, sh, st)))

    def sync(self):
        check(_lib().b200h_sync(self.h))

    def flush(self):
        """Plan and launch everything pending without waiting for the device."""
        check(_lib().b200h_flush(self.h))

    def cache_stats(self) -> CacheStats:
        cs = CacheStats()
        check(_lib().b200h_cache_stats_get(self.h, C.byref(cs)))
        return cs

    def clear_cache(self):
        check(_lib().b200h_cache_clear(self.h))

    def fuser(self, kind: int) -> "Fuser":
        h = C.c_void_p()
        check(_lib().b200h_fuser_create(self.h, kind, C.byref(h)))
        return Fuser(self, h)

    def blocks(self):
        out = []
        for i in range(_lib().b200h_block_count(self.h)):
            b = BlockInfo()
            check(_lib().b200h_block_get(self.h, i, C.byref(b)))
            out.append(b)
        return out

    def clear_blocks(self):
        check(_lib().b200h_block_clear(self.h))


class Fuser:
    """`OperationFuser` (crates/burn-fusion/src/backend.rs:187-206) over the stream's pending queue."""

    def __init__(self, stream: FusionStream, h):
        self.stream, self.h = stream, h

    def __del__(self):
        if getattr(self, "h", None):
            _lib().b200h_fuser_destroy(self.h)
            self.h = None

    def fuse_next(self): check(_lib().b200h_fuser_fuse_next(self.h))
    @property
    def status(self): return _lib().b200h_fuser_status_get(self.h)
    def __len__(self): return _lib().b200h_fuser_len(self.h)
    def reset(self): check(_lib().b200h_fuser_reset(self.h))

    def properties(self):
        score, ready = C.c_uint64(), C.c_int32()
        check(_lib().b200h_fuser_properties(self.h, C.byref(score), C.byref(ready)))
        return score.value, bool(ready.value)

    def clone_dyn(self) -> "Fuser":
        h = C.c_void_p()
        check(_lib().b200h_fuser_clone(self.h, C.byref(h)))
        return Fuser(self.stream, h)

    def fuse_until_closed(self):
        while self.status == FUSER_OPEN:
            try:
                self.fuse_next()
            except abi.B200Error:
                break                       # queue exhausted while still open
        return self

    def finish(self) -> "Optimization":
        h = C.c_void_p()
        check(_lib().b200h_fuser_finish(self.h, C.byref(h)))
        return Optimization(h)


class Optimization:
    """`Optimization` (backend.rs:226-234): execute on a stream's queue head, to_state / from_state."""

    def __init__(self, h):
        self.h = h

    def __del__(self):
        if getattr(self, "h", None):
            _lib().b200h_optimization_destroy(self.h)
            self.h = None

    def __len__(self): return _lib().b200h_optimization_len(self.h)
    @property
    def name(self): return _lib().b200h_optimization_name(self.h).decode()
    def execute(self, stream: FusionStream): check(_lib().b200h_optimization_execute(self.h, stream.h))

    def to_state(self) -> bytes:
        n = C.c_uint64()
        check(_lib().b200h_optimization_to_state(self.h, None, 0, C.byref(n)))
        buf = C.create_string_buffer(n.value)
        check(_lib().b200h_optimization_to_state(self.h, buf, n.value, C.byref(n)))
        return buf.raw[:n.value]

    @staticmethod
    def from_state(state: bytes) -> "Optimization":
        h = C.c_void_p()
        check(_lib().b200h_optimization_from_state(state, len(state), C.byref(h)))
        return Optimization(h)


def gelu(x: LazyTensor) -> LazyTensor:
    """ActivationOps::gelu default: the five primitive ops, each intermediate dropped after use
    (crates/burn-backend/src/backend/ops/activation.rs:69-76)."""
    t = x.div_scalar(1.4142135623730951)
    e = t.erf(); t.drop()
    p = e.add_scalar(1.0); e.drop()
    m = x.mul(p); p.drop()
    y = m.div_scalar(2.0); m.drop()
    return y


def softmax(x: LazyTensor, dim: int) -> LazyTensor:
    """The op chain ActivationOps::softmax records (activation.rs:250-256), intermediates dropped as
    they go out of scope; along the last axis the ReduceBroadcasted fuser turns it into one kernel."""
    m = x.max_dim(dim)
    d = x.sub(m)
    m.drop()
    e = d.exp()
    d.drop()
    s = e.sum_dim(dim)
    y = e.div(s)
    e.drop()
    s.drop()
    return y


def log_softmax(x: LazyTensor, dim: int) -> LazyTensor:
    """ActivationOps::log_softmax (activation.rs:271-276)."""
    m = x.max_dim(dim)
    d = x.sub(m)
    m.drop()
    e = d.exp()
    s = e.sum_dim(dim)
    e.drop()
    ls = s.log()
    s.drop()
    y = d.sub(ls)
    d.drop()
    ls.drop()
    return y


def layer_norm(x: LazyTensor, gamma: LazyTensor, beta: LazyTensor | None, eps: float) -> LazyTensor:
    """ModuleOps::layer_norm's default chain (ops/modules/base.rs:846-877) along the last axis; gamma / beta are
    [1, …, d_model] tensors (the reference reshapes them, a metadata op)."""
    dim = len(x.shape) - 1
    mean = x.mean_dim(dim)
    c = x.sub(mean)
    mean.drop()
    sq = c.mul(c)
    var = sq.mean_dim(dim)
    sq.drop()
    ve = var.add_scalar(eps)
    var.drop()
    den = ve.sqrt()
    ve.drop()
    n = c.div(den)
    c.drop()
    den.drop()
    y = n.mul(gamma)
    n.drop()
    if beta is not None:
        z = y.add(beta)
        y.drop()
        return z
    return y
