"""Thin Python binding over the C ABI: device tensors, views and tape building.

This is the moral equivalent of `CubeTensor` (crates/burn-cubecl/src/tensor/base.rs:20-33):
a device handle plus shape/strides metadata, where swap_dims / permute / expand /
reshape-of-contiguous are metadata-only views (crates/burn-cubecl/src/ops/base.rs:137-139,
213-240).  All compute goes through libburn_b200.so; nothing here touches torch
or falls back to numpy for arithmetic.
"""
from __future__ import annotations

import ctypes as C
import struct
from typing import Iterable, Sequence

import numpy as np

from . import _abi as abi
from ._abi import check

_NP2DT = {
    np.dtype(np.float32): abi.F32, np.dtype(np.float16): abi.F16, np.dtype(np.int32): abi.I32,
    np.dtype(np.int64): abi.I64, np.dtype(np.bool_): abi.BOOL, np.dtype(np.uint8): abi.U8,
}
_DT2NP = {abi.F32: np.float32, abi.F16: np.float16, abi.BF16: np.uint16, abi.I32: np.int32,
          abi.I64: np.int64, abi.BOOL: np.bool_, abi.U8: np.uint8}

_initialized = False


def init(device: int = 0) -> None:
    global _initialized
    check(abi.load().b200_init(device))
    _initialized = True


def _lib():
    if not _initialized:
        init(0)
    return abi.load()


def sync(stream=None) -> None:
    check(_lib().b200_stream_sync(stream))


class Graph:
    """A captured launch sequence (b200_graph_*): `with Graph.capture() as g: ...` records every
    launch / alloc / free issued on the default stream; `g.launch()` replays it."""

    def __init__(self):
        self.handle = C.c_void_p()
        self.kernel_nodes = self.total_nodes = 0
        self._stream = None

    @classmethod
    def capture(cls, stream=None) -> "Graph":
        g = cls()
        g._stream = stream
        return g

    def __enter__(self):
        import gc
        gc.collect()            # no stray pre-capture buffers may be freed inside the capture
        check(_lib().b200_graph_begin(self._stream))
        return self

    def __exit__(self, et, ev, tb):
        import gc
        gc.collect()            # buffers allocated inside the capture are freed inside it
        st = _lib().b200_graph_end(self._stream, C.byref(self.handle))
        if et is None:
            check(st)
            k, t = C.c_uint64(), C.c_uint64()
            check(_lib().b200_graph_node_count(self.handle, C.byref(k), C.byref(t)))
            self.kernel_nodes, self.total_nodes = k.value, t.value
        return False

    def launch(self, stream=None) -> None:
        check(_lib().b200_graph_launch(self.handle, stream))

    def destroy(self) -> None:
        """Must run before b200_comm_destroy of any communicator whose collectives were captured:
        ncclCommDestroy waits for every graph that still references the communicator."""
        if self.handle:
            check(_lib().b200_graph_destroy(self.handle))
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class Storage:
    """Refcounted device allocation (freed stream-ordered when the last view dies)."""

    def __init__(self, nbytes: int, stream=None):
        self.ptr = C.c_void_p()
        self.nbytes = nbytes
        self.stream = stream
        check(_lib().b200_alloc(C.byref(self.ptr), nbytes, stream))

    def __del__(self):
        try:
            if self.ptr:
                abi.load().b200_free(self.ptr, self.stream)
        except Exception:
            pass


def contiguous_strides(shape: Sequence[int]) -> tuple[int, ...]:
    st, acc = [], 1
    for s in reversed(shape):
        st.append(acc)
        acc *= max(int(s), 1)
    return tuple(reversed(st))


class DeviceTensor:
    def __init__(self, storage: Storage, dtype: int, shape, strides=None, offset: int = 0):
        self.storage = storage
        self.dtype = dtype
        self.shape = tuple(int(s) for s in shape)
        self.strides = tuple(strides) if strides is not None else contiguous_strides(self.shape)
        self.offset = offset  # in elements

    # ---- construction
    @staticmethod
    def empty(shape, dtype: int = abi.F32, stream=None) -> "DeviceTensor":
        n = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
        return DeviceTensor(Storage(max(n, 1) * abi.DTYPE_SIZE[dtype], stream), dtype, shape)

    @staticmethod
    def from_numpy(a: np.ndarray, stream=None, dtype: int | None = None) -> "DeviceTensor":
        a = np.ascontiguousarray(a)
        if dtype is None:
            dtype = _NP2DT[a.dtype]
        if dtype == abi.BOOL:
            a = a.astype(np.uint8)
        t = DeviceTensor.empty(a.shape, dtype, stream)
        check(_lib().b200_memcpy_h2d(t.storage.ptr, a.ctypes.data_as(C.c_void_p), a.nbytes, stream))
        check(_lib().b200_stream_sync(stream))  # `a` may be a temporary
        return t

    @staticmethod
    def from_bf16_of(a: np.ndarray, stream=None) -> "DeviceTensor":
        """Uploads f32 data rounded (RN-even) to bf16 storage."""
        bits = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
        rounded = ((bits + 0x7FFF + ((bits >> 16) & 1)) >> 16).astype(np.uint16)
        return DeviceTensor.from_numpy(rounded, stream, dtype=abi.BF16)

    # ---- readback (float_into_data)
    def numpy(self, stream=None) -> np.ndarray:
        t = self if self.is_contiguous() and self.offset == 0 else self.contiguous(stream)
        out = np.empty(t.shape, dtype=_DT2NP[t.dtype] if t.dtype != abi.BOOL else np.uint8)
        if out.nbytes:
            check(_lib().b200_memcpy_d2h(out.ctypes.data_as(C.c_void_p), t.data_ptr(), out.nbytes, stream))
        check(_lib().b200_stream_sync(stream))
        if t.dtype == abi.BOOL:
            out = out.astype(np.bool_)
        if t.dtype == abi.BF16:
            out = (out.astype(np.uint32) << 16).view(np.float32)
        return out

    # ---- metadata
    @property
    def ndim(self) -> int:
        return len(self.shape)

    @property
    def numel(self) -> int:
        return int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1

    def data_ptr(self) -> int:
        return (self.storage.ptr.value or 0) + self.offset * abi.DTYPE_SIZE[self.dtype]

    def is_contiguous(self) -> bool:
        expect = 1
        for s, st in zip(reversed(self.shape), reversed(self.strides)):
            if s != 1 and st != expect:
                return False
            expect *= s
        return True

    def desc(self) -> abi.Tensor:
        d = abi.Tensor()
        d.ptr = self.data_ptr()
        d.dtype = self.dtype
        d.rank = len(self.shape)
        for i, (s, st) in enumerate(zip(self.shape, self.strides)):
            d.shape[i] = s
            d.strides[i] = st
        return d

    # ---- views (metadata only)
    def swap_dims(self, d0: int, d1: int) -> "DeviceTensor":
        sh, st = list(self.shape), list(self.strides)
        sh[d0], sh[d1] = sh[d1], sh[d0]
        st[d0], st[d1] = st[d1], st[d0]
        return DeviceTensor(self.storage, self.dtype, sh, st, self.offset)

    def permute(self, axes: Sequence[int]) -> "DeviceTensor":
        return DeviceTensor(self.storage, self.dtype, [self.shape[a] for a in axes],
                            [self.strides[a] for a in axes], self.offset)

    def expand(self, shape: Sequence[int]) -> "DeviceTensor":
        pad = len(shape) - len(self.shape)
        sh = (1,) * pad + self.shape
        st = (0,) * pad + self.strides
        new_st = []
        for s_old, s_new, stride in zip(sh, shape, st):
            if s_old == s_new:
                new_st.append(stride)
            elif s_old == 1:
                new_st.append(0)
            else:
                raise ValueError(f"cannot expand {self.shape} to {tuple(shape)}")
        return DeviceTensor(self.storage, self.dtype, shape, new_st, self.offset)

    def reshape(self, shape: Sequence[int], stream=None) -> "DeviceTensor":
        t = self if self.is_contiguous() else self.contiguous(stream)
        return DeviceTensor(t.storage, t.dtype, shape, None, t.offset)

    def slice(self, ranges: Sequence[tuple[int, int]]) -> "DeviceTensor":
        off = self.offset
        sh = list(self.shape)
        for d, (lo, hi) in enumerate(ranges):
            off += lo * self.strides[d]
            sh[d] = hi - lo
        return DeviceTensor(self.storage, self.dtype, sh, self.strides, off)

    def contiguous(self, stream=None) -> "DeviceTensor":
        out = DeviceTensor.empty(self.shape, self.dtype, stream)
        if self.numel:
            s, d = self.desc(), out.desc()
            check(_lib().b200_launch_copy(C.byref(s), C.byref(d), stream))
        return out


# ---------------------------------------------------------------- tapes
def f32_bits(x: float) -> int:
    return struct.unpack("<I", struct.pack("<f", float(np.float32(x))))[0]


def i32_bits(x: int) -> int:
    return int(x) & 0xFFFFFFFF


class TapeBuilder:
    """Assembles a b200_tape.  Operands: 'acc', ('in', i), ('tmp', i), ('f', value), ('i', value)."""

    def __init__(self):
        self.ops: list[tuple[int, int, int, int, int, int]] = []
        self.scalars: list[int] = []

    def _arg(self, a) -> int:
        if a is None or a == "acc":
            return abi.ARG_ACC
        kind, v = a
        if kind == "in":
            return abi.ARG_INPUT(v)
        if kind == "tmp":
            return abi.ARG_TEMP(v)
        bits = f32_bits(v) if kind == "f" else i32_bits(v)
        if bits in self.scalars:
            return abi.ARG_SCALAR(self.scalars.index(bits))
        self.scalars.append(bits)
        return abi.ARG_SCALAR(len(self.scalars) - 1)

    def op(self, name: str, a=None, b=None, c=None, tmp: int | None = None, out: int | None = None) -> "TapeBuilder":
        self.ops.append((abi.OP[name], self._arg(a), self._arg(b) if b is not None else 0,
                         self._arg(c) if c is not None else 0,
                         abi.DST_NONE if tmp is None else tmp, abi.DST_NONE if out is None else out))
        return self

    def build(self):
        n = len(self.ops)
        ops = (abi.TapeOp * max(n, 1))()
        for i, (o, a, b, c, t, out) in enumerate(self.ops):
            ops[i].op, ops[i].a, ops[i].b, ops[i].c, ops[i].dst_temp, ops[i].dst_out = o, a, b, c, t, out
        sc = (C.c_uint32 * max(len(self.scalars), 1))(*self.scalars)
        tape = abi.Tape()
        tape.ops = C.cast(ops, C.POINTER(abi.TapeOp))
        tape.n_ops = n
        tape.scalars = C.cast(sc, C.POINTER(C.c_uint32))
        tape.n_scalars = len(self.scalars)
        tape._keep = (ops, sc)  # keep the buffers alive with the struct
        return tape


def _descs(ts: Iterable[DeviceTensor]):
    ts = list(ts)
    arr = (abi.Tensor * max(len(ts), 1))()
    for i, t in enumerate(ts):
        arr[i] = t.desc()
    return arr, len(ts)


def launch_elemwise(tape, inputs: Sequence[DeviceTensor], outputs: Sequence[DeviceTensor], ref_shape,
                    stream=None) -> None:
    ins, n_in = _descs(inputs)
    outs, n_out = _descs(outputs)
    shape = (C.c_int64 * max(len(ref_shape), 1))(*ref_shape)
    check(_lib().b200_launch_elemwise(C.byref(tape), ins, n_in, outs, n_out, len(ref_shape), shape, stream))


def launch_reduce(kind: int, axis: int, in_shape, inputs, outputs, read=None, write=None, write_inputs=(),
                  stream=None) -> None:
    ins, n_in = _descs(inputs)
    wins, n_win = _descs(write_inputs)
    outs, n_out = _descs(outputs)
    shape = (C.c_int64 * max(len(in_shape), 1))(*in_shape)
    check(_lib().b200_launch_reduce(kind, axis, len(in_shape), shape,
                                    C.byref(read) if read is not None else None, ins, n_in,
                                    C.byref(write) if write is not None else None, wins, n_win, outs, n_out,
                                    stream))


def launch_reduce_full(kind: int, x: DeviceTensor, out: DeviceTensor, stream=None) -> None:
    a, b = x.desc(), out.desc()
    check(_lib().b200_launch_reduce_full(kind, C.byref(a), C.byref(b), stream))
