"""Eager op layer over the C ABI, named after the reference's backend traits.

Each function is the B200 counterpart of the `FloatTensorOps` / `ActivationOps` entry point of
the same name (crates/burn-backend/src/backend/ops/tensor.rs, ops/activation.rs): same argument
meaning, keepdim reductions, broadcast rules, and errors raised where the reference panics.
Composite activations are launched as ONE fused tape of the primitive ops the reference's
default implementations issue (activation.rs:37-76,250-276) — i.e. what burn-fusion's
ElementWise / Reduce fusers hand to `Optimization::execute`.

All arithmetic runs in libburn_b200.so; nothing here computes on the host.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _abi as abi
from . import device as dv
from ._abi import check
from .device import DeviceTensor, TapeBuilder

SQRT_2 = 1.4142135623730951


class ShapeError(ValueError):
    """Raised where the reference panics on incompatible shapes / dims."""


def _bshape(a: Sequence[int], b: Sequence[int]):
    if len(a) != len(b):
        raise ShapeError(f"rank mismatch: {tuple(a)} vs {tuple(b)}")
    out = []
    for x, y in zip(a, b):
        if x != y and x != 1 and y != 1:
            raise ShapeError(f"shapes {tuple(a)} and {tuple(b)} are not broadcastable")
        out.append(max(x, y))
    return tuple(out)


def _launch(tb: TapeBuilder, inputs, shape, out_dtype=abi.F32) -> DeviceTensor:
    out = DeviceTensor.empty(shape, out_dtype)
    if out.numel:
        dv.launch_elemwise(tb.build(), inputs, [out], shape)
    return out


def _binary(opname: str, lhs: DeviceTensor, rhs: DeviceTensor, out_dtype=abi.F32) -> DeviceTensor:
    shape = _bshape(lhs.shape, rhs.shape)
    tb = TapeBuilder().op(opname, ("in", 0), ("in", 1), out=0)
    return _launch(tb, [lhs.expand(shape), rhs.expand(shape)], shape, out_dtype)


def _scalar(opname: str, lhs: DeviceTensor, s, out_dtype=abi.F32, kind="f") -> DeviceTensor:
    tb = TapeBuilder().op(opname, ("in", 0), (kind, s), out=0)
    return _launch(tb, [lhs], lhs.shape, out_dtype)


def _unary(opname: str, x: DeviceTensor, out_dtype=abi.F32) -> DeviceTensor:
    return _launch(TapeBuilder().op(opname, ("in", 0), out=0), [x], x.shape, out_dtype)


# ---- FloatTensorOps: arithmetic (tensor.rs:183-329)
def float_add(a, b): return _binary("ADD_F", a, b)
def float_sub(a, b): return _binary("SUB_F", a, b)
def float_mul(a, b): return _binary("MUL_F", a, b)
def float_div(a, b): return _binary("DIV_F", a, b)
def float_remainder(a, b): return _binary("REMT_F", a, b)   # tensor-tensor form: a - b*floor(a/b) in f64
def float_powf(a, b): return _binary("POW_F", a, b)
def float_add_scalar(a, s): return _scalar("ADD_F", a, s)
def float_sub_scalar(a, s): return _scalar("SUB_F", a, s)
def float_mul_scalar(a, s): return _scalar("MUL_F", a, s)
def float_div_scalar(a, s): return _scalar("DIV_F", a, s)
def float_remainder_scalar(a, s): return _scalar("REM_F", a, s)
def float_powf_scalar(a, s): return _scalar("POW_F", a, s)

# ---- unary math (tensor.rs:1024-1442)
def float_exp(a): return _unary("EXP_F", a)
def float_log(a): return _unary("LOG_F", a)
def float_log1p(a): return _unary("LOG1P_F", a)
def float_sqrt(a): return _unary("SQRT_F", a)
def float_abs(a): return _unary("ABS_F", a)
def float_neg(a): return _unary("NEG_F", a)
def float_recip(a): return _unary("RECIP_F", a)
def float_tanh(a): return _unary("TANH_F", a)
def float_erf(a): return _unary("ERF_F", a)
def float_sin(a): return _unary("SIN_F", a)
def float_cos(a): return _unary("COS_F", a)
def float_floor(a): return _unary("FLOOR_F", a)
def float_ceil(a): return _unary("CEIL_F", a)
def float_round(a): return _unary("ROUND_F", a)
def float_trunc(a): return _unary("TRUNC_F", a)
def float_sign(a): return _unary("SIGN_F", a)
def float_tan(a): return _unary("TAN_F", a)


def float_powi_scalar(a, s):
    """float_powi_scalar (crates/burn-backend/src/backend/ops/tensor.rs:1095-1104): small exponents are products /
    reciprocals, everything else defers to powf_scalar."""
    s = int(s)
    if s == 0:
        return _launch(TapeBuilder().op("MOV", ("f", 1.0), out=0), [a], a.shape)      # float_ones
    if s == 1:
        return a
    if s == 2:
        return float_mul(a, a)
    if s == -1:
        return float_recip(a)
    if s == -2:
        return float_recip(float_mul(a, a))
    return float_powf_scalar(a, float(s))


def float_clamp_min(a, lo):
    """float_clamp_min default (tensor.rs:207-214): mask_fill(x, x < min, min)."""
    return float_mask_fill(a, float_lower_elem(a, lo), lo)


def float_clamp_max(a, hi):
    """float_clamp_max default (tensor.rs:223-230): mask_fill(x, x > max, max)."""
    return float_mask_fill(a, float_greater_elem(a, hi), hi)


def float_clamp(a, lo, hi):
    return _launch(TapeBuilder().op("CLAMP_F", ("in", 0), ("f", lo), ("f", hi), out=0), [a], a.shape)


# ---- comparisons → bool (tensor.rs:675-850)
def float_equal(a, b): return _binary("EQ_F", a, b, abi.BOOL)
def float_not_equal(a, b): return _binary("NE_F", a, b, abi.BOOL)
def float_not_equal_elem(a, s): return _scalar("NE_F", a, s, abi.BOOL)
def float_is_nan(a): return _unary("ISNAN_F", a, abi.BOOL)
def float_is_inf(a): return _unary("ISINF_F", a, abi.BOOL)
def float_greater(a, b): return _binary("GT_F", a, b, abi.BOOL)
def float_greater_equal(a, b): return _binary("GE_F", a, b, abi.BOOL)
def float_lower(a, b): return _binary("LT_F", a, b, abi.BOOL)
def float_lower_equal(a, b): return _binary("LE_F", a, b, abi.BOOL)
def float_equal_elem(a, s): return _scalar("EQ_F", a, s, abi.BOOL)
def float_greater_elem(a, s): return _scalar("GT_F", a, s, abi.BOOL)
def float_greater_equal_elem(a, s): return _scalar("GE_F", a, s, abi.BOOL)
def float_lower_elem(a, s): return _scalar("LT_F", a, s, abi.BOOL)
def float_lower_equal_elem(a, s): return _scalar("LE_F", a, s, abi.BOOL)


# ---- masks (tensor.rs:609-630)
def float_mask_fill(x: DeviceTensor, mask: DeviceTensor, value: float) -> DeviceTensor:
    shape = _bshape(x.shape, mask.shape)
    tb = TapeBuilder().op("SELECT", ("in", 0), ("f", value), ("in", 1), out=0)
    return _launch(tb, [x.expand(shape), mask.expand(shape)], shape)


def float_mask_where(x: DeviceTensor, mask: DeviceTensor, source: DeviceTensor) -> DeviceTensor:
    shape = _bshape(_bshape(x.shape, mask.shape), source.shape)
    tb = TapeBuilder().op("SELECT", ("in", 0), ("in", 1), ("in", 2), out=0)
    return _launch(tb, [x.expand(shape), source.expand(shape), mask.expand(shape)], shape)


# ---- casts (tensor.rs:135,1013)
def float_into_int(x: DeviceTensor, dtype=abi.I32) -> DeviceTensor:
    return _unary("F2I", x, dtype)


def float_cast(x: DeviceTensor, dtype: int) -> DeviceTensor:
    return _unary("MOV", x, dtype)


def int_into_float(x: DeviceTensor, dtype=abi.F32) -> DeviceTensor:
    """int_into_float / bool_into_float (int_tensor.rs:119, bool_tensor.rs:102)."""
    return _unary("B2F" if x.dtype in (abi.BOOL, abi.U8) else "I2F", x, dtype)


# ---- reductions (tensor.rs:883-949,1484-1636)
def _check_dim(x: DeviceTensor, dim: int) -> int:
    if dim < 0:
        dim += x.ndim
    if not 0 <= dim < x.ndim:
        raise ShapeError(f"dim {dim} out of range for rank {x.ndim}")
    return dim


def _reduce(kind: int, x: DeviceTensor, dim: int, out_dtype=abi.F32) -> DeviceTensor:
    dim = _check_dim(x, dim)
    shape = list(x.shape)
    shape[dim] = 1
    out = DeviceTensor.empty(shape, out_dtype)
    if out.numel:
        dv.launch_reduce(kind, dim, x.shape, [x], [out])
    return out


def float_sum_dim(x, dim): return _reduce(abi.RED_SUM, x, dim)
def float_mean_dim(x, dim): return _reduce(abi.RED_MEAN, x, dim)
def float_prod_dim(x, dim): return _reduce(abi.RED_PROD, x, dim)
def float_max_dim(x, dim): return _reduce(abi.RED_MAX, x, dim)
def float_min_dim(x, dim): return _reduce(abi.RED_MIN, x, dim)
def float_argmax(x, dim, out_dtype=abi.I32): return _reduce(abi.RED_ARGMAX, x, dim, out_dtype)
def float_argmin(x, dim, out_dtype=abi.I32): return _reduce(abi.RED_ARGMIN, x, dim, out_dtype)


def _reduce_full(kind: int, x: DeviceTensor) -> DeviceTensor:
    out = DeviceTensor.empty((1,))
    dv.launch_reduce_full(kind, x, out)
    return out


def float_sum(x): return _reduce_full(abi.RED_SUM, x)
def float_mean(x): return _reduce_full(abi.RED_MEAN, x)
def float_max(x): return _reduce_full(abi.RED_MAX, x)
def float_min(x): return _reduce_full(abi.RED_MIN, x)
def float_prod(x): return _reduce_full(abi.RED_PROD, x)
def float_max_abs(x): return _reduce_full(abi.RED_MAXABS, x)
def float_max_abs_dim(x, dim): return _reduce(abi.RED_MAXABS, x, dim)


# ---- matmul (tensor.rs:341)
def float_matmul(lhs: DeviceTensor, rhs: DeviceTensor, precision: int = abi.MM_F32X3, epilogue=None,
                 epi_inputs: Sequence[DeviceTensor] = ()) -> DeviceTensor:
    if lhs.ndim != rhs.ndim or lhs.ndim < 2:
        raise ShapeError("matmul operands must have the same rank >= 2")
    if lhs.shape[-1] != rhs.shape[-2]:
        raise ShapeError(f"matmul inner dims differ: {lhs.shape} x {rhs.shape}")
    batch = _bshape(lhs.shape[:-2], rhs.shape[:-2])
    out = DeviceTensor.empty(tuple(batch) + (lhs.shape[-2], rhs.shape[-1]))
    a, b, c = lhs.desc(), rhs.desc(), out.desc()
    ws_bytes = C.c_uint64()
    check(abi.load().b200_matmul_workspace_bytes(C.byref(a), C.byref(b), precision, C.byref(ws_bytes)))
    ws = dv.Storage(ws_bytes.value) if ws_bytes.value else None
    epi, n_epi = dv._descs(epi_inputs)
    check(abi.load().b200_launch_matmul(C.byref(a), C.byref(b), C.byref(c), precision,
                                        C.byref(epilogue) if epilogue is not None else None, epi, n_epi,
                                        ws.ptr if ws else None, ws_bytes.value, None))
    return out


# ---- indexing (tensor.rs:435-543)
def float_gather(dim: int, x: DeviceTensor, indices: DeviceTensor) -> DeviceTensor:
    out = DeviceTensor.empty(indices.shape, x.dtype)
    a, b, c = x.desc(), indices.desc(), out.desc()
    check(abi.load().b200_launch_gather(dim, C.byref(a), C.byref(b), C.byref(c), None))
    return out


def float_scatter_add(dim: int, x: DeviceTensor, indices: DeviceTensor, value: DeviceTensor) -> DeviceTensor:
    out = x.contiguous()  # the reference returns a new tensor unless it owns `x`
    a, b, c = out.desc(), indices.desc(), value.desc()
    check(abi.load().b200_launch_scatter_add(dim, C.byref(a), C.byref(b), C.byref(c), None))
    return out


def float_select(x: DeviceTensor, dim: int, indices: DeviceTensor) -> DeviceTensor:
    dim = _check_dim(x, dim)
    shape = list(x.shape)
    shape[dim] = indices.shape[0]
    out = DeviceTensor.empty(shape, x.dtype)
    a, b, c = x.desc(), indices.desc(), out.desc()
    check(abi.load().b200_launch_select(dim, C.byref(a), C.byref(b), C.byref(c), None))
    return out


def float_select_add(x: DeviceTensor, dim: int, indices: DeviceTensor, value: DeviceTensor) -> DeviceTensor:
    dim = _check_dim(x, dim)
    out = x.contiguous()
    a, b, c = out.desc(), indices.desc(), value.desc()
    check(abi.load().b200_launch_select_add(dim, C.byref(a), C.byref(b), C.byref(c), None))
    return out


# ---- data movement (tensor.rs:161,410,592,1460)
def _clone(x: DeviceTensor) -> DeviceTensor:
    out = DeviceTensor.empty(x.shape, x.dtype)
    if out.numel:
        a, b = x.desc(), out.desc()
        check(abi.load().b200_launch_copy(C.byref(a), C.byref(b), None))
    return out


def _canon_ranges(shape, ranges):
    """Tensor::slice canonicalisation (crates/burn-tensor/src/tensor/api/base.rs `slice`): negative bounds count from
    the end, ends clamp to the dimension, unspecified trailing dims are full, a descending range is empty."""
    full = []
    for i, n in enumerate(shape):
        if i < len(ranges):
            lo, hi = ranges[i]
            lo = lo + n if lo < 0 else lo
            hi = n if hi is None else (hi + n if hi < 0 else hi)
            lo, hi = min(max(lo, 0), n), min(max(hi, 0), n)
            full.append((lo, max(hi, lo)))
        else:
            full.append((0, n))
    return full


def float_slice(x: DeviceTensor, ranges) -> DeviceTensor:
    """float_slice (tensor.rs:580): a strided view — no launch; consumers read it in place."""
    return x.slice(_canon_ranges(x.shape, ranges))


def float_slice_assign(x: DeviceTensor, ranges, value: DeviceTensor) -> DeviceTensor:
    """x[ranges] = value on a copy (the reference mutates in place only when it owns `x`)."""
    out = _clone(x)
    full = _canon_ranges(x.shape, ranges)
    starts = (C.c_int64 * x.ndim)(*[r[0] for r in full])
    ends = (C.c_int64 * x.ndim)(*[r[1] for r in full])
    a, b = out.desc(), value.desc()
    check(abi.load().b200_launch_slice_assign(C.byref(a), starts, ends, C.byref(b), None))
    return out


def float_cat(tensors: Sequence[DeviceTensor], dim: int) -> DeviceTensor:
    dim = _check_dim(tensors[0], dim)
    shape = list(next((t for t in tensors if t.numel), tensors[0]).shape)
    shape[dim] = sum(t.shape[dim] for t in tensors if t.ndim == len(shape))
    out = DeviceTensor.empty(shape, tensors[0].dtype)
    descs = (abi.Tensor * len(tensors))(*[t.desc() for t in tensors])
    od = out.desc()
    check(abi.load().b200_launch_cat(descs, len(tensors), dim, C.byref(od), None))
    return out


def float_repeat_dim(x: DeviceTensor, dim: int, times: int) -> DeviceTensor:
    dim = _check_dim(x, dim)
    shape = list(x.shape)
    shape[dim] *= times
    out = DeviceTensor.empty(shape, x.dtype)
    a, b = x.desc(), out.desc()
    check(abi.load().b200_launch_repeat_dim(C.byref(a), dim, times, C.byref(b), None))
    return out


def float_flip(x: DeviceTensor, axes: Sequence[int]) -> DeviceTensor:
    out = DeviceTensor.empty(x.shape, x.dtype)
    ax = (C.c_int32 * max(len(axes), 1))(*[_check_dim(x, a) for a in axes])
    a, b = x.desc(), out.desc()
    check(abi.load().b200_launch_flip(C.byref(a), ax, len(axes), C.byref(b), None))
    return out


# ---- ActivationOps defaults, fused (activation.rs:37-76,139-160,250-276)
def relu(x: DeviceTensor) -> DeviceTensor:
    tb = (TapeBuilder().op("LE_F", ("in", 0), ("f", 0.0), tmp=0)
          .op("SELECT", ("in", 0), ("f", 0.0), ("tmp", 0), out=0))
    return _launch(tb, [x], x.shape)


def gelu(x: DeviceTensor) -> DeviceTensor:
    tb = (TapeBuilder().op("DIV_F", ("in", 0), ("f", SQRT_2)).op("ERF_F", "acc").op("ADD_F", "acc", ("f", 1.0))
          .op("MUL_F", ("in", 0), "acc").op("DIV_F", "acc", ("f", 2.0), out=0))
    return _launch(tb, [x], x.shape)


def sigmoid(x: DeviceTensor) -> DeviceTensor:
    return _unary("SIGMOID_F", x)


def softmax(x: DeviceTensor, dim: int) -> DeviceTensor:
    """max_dim → sub → exp → sum_dim → div: two reduce launches with fused read/write tapes."""
    dim = _check_dim(x, dim)
    m = float_max_dim(x, dim)
    # e = exp(x - max) written once, while its row sum is reduced by the same launch?  The reduce
    # entry point has no elementwise outputs, so: fused exp(x-max) → sum, then one fused divide.
    shape = x.shape
    red = list(shape)
    red[dim] = 1
    s = DeviceTensor.empty(red)
    rd = TapeBuilder().op("SUB_F", ("in", 0), ("in", 1)).op("EXP_F", "acc")
    dv.launch_reduce(abi.RED_SUM, dim, shape, [x, m.expand(shape)], [s], read=rd.build())
    tb = (TapeBuilder().op("SUB_F", ("in", 0), ("in", 1)).op("EXP_F", "acc")
          .op("DIV_F", "acc", ("in", 2), out=0))
    return _launch(tb, [x, m.expand(shape), s.expand(shape)], shape)


def log_softmax(x: DeviceTensor, dim: int) -> DeviceTensor:
    dim = _check_dim(x, dim)
    m = float_max_dim(x, dim)
    shape = x.shape
    red = list(shape)
    red[dim] = 1
    lse = DeviceTensor.empty(red)
    rd = TapeBuilder().op("SUB_F", ("in", 0), ("in", 1)).op("EXP_F", "acc")
    wr = TapeBuilder().op("LOG_F", ("in", 0), out=0)
    dv.launch_reduce(abi.RED_SUM, dim, shape, [x, m.expand(shape)], [lse], read=rd.build(), write=wr.build())
    tb = TapeBuilder().op("SUB_F", ("in", 0), ("in", 1)).op("SUB_F", "acc", ("in", 2), out=0)
    return _launch(tb, [x, m.expand(shape), lse.expand(shape)], shape)


# ---- row-resident fused kernels (ReduceBroadcasted analogue)
def softmax_rows(x: DeviceTensor, log: bool = False, out: DeviceTensor | None = None) -> DeviceTensor:
    """softmax / log_softmax along the LAST axis in one kernel (b200_launch_softmax); `out` may alias `x`
    (each row is read completely before it is written)."""
    out = DeviceTensor.empty(x.shape) if out is None else out
    a, b = x.desc(), out.desc()
    check(abi.load().b200_launch_softmax(C.byref(a), C.byref(b), 1 if log else 0, None))
    return out


def layer_norm(x: DeviceTensor, gamma: DeviceTensor | None, beta: DeviceTensor | None, eps: float) -> DeviceTensor:
    """ModuleOps::layer_norm over the last axis in one kernel (b200_launch_layer_norm)."""
    out = DeviceTensor.empty(x.shape)
    a, o = x.desc(), out.desc()
    g = gamma.desc() if gamma is not None else None
    b = beta.desc() if beta is not None else None
    check(abi.load().b200_launch_layer_norm(C.byref(a), C.byref(g) if g is not None else None,
                                            C.byref(b) if b is not None else None, float(eps), C.byref(o), None))
    return out


def softmax_backward(y: DeviceTensor, dy: DeviceTensor, mask: DeviceTensor | None = None, div: float = 1.0,
                     out: DeviceTensor | None = None) -> DeviceTensor:
    """dx = (dy - sum(dy*y, -1)) * y / div, 0 where `mask` (b200_launch_softmax_backward); `out` may alias `dy`."""
    out = DeviceTensor.empty(y.shape) if out is None else out
    a, g, o = y.desc(), dy.desc(), out.desc()
    m = mask.desc() if mask is not None else None
    check(abi.load().b200_launch_softmax_backward(C.byref(a), C.byref(g), C.byref(m) if m is not None else None,
                                                  float(div), C.byref(o), None))
    return out


def layer_norm_backward(x: DeviceTensor, dy: DeviceTensor, gamma: DeviceTensor | None, eps: float, want_dx_sum: bool = False):
    """(dx, dgamma, dbeta) of layer_norm over the last axis: one row-resident kernel plus the
    deterministic column reduce of its per-CTA partials (b200_launch_layer_norm_backward).
    want_dx_sum: also return colsum(dx) — the bias gradient of the Linear that produced x — from a third partial the
    same kernel emits (b200_launch_layer_norm_backward_ex), instead of a separate pass over dx."""
    lib = abi.load()
    d = x.shape[-1]
    dx = DeviceTensor.empty(x.shape)
    a, g, o = x.desc(), dy.desc(), dx.desc()
    n = C.c_int32()
    check(lib.b200_layer_norm_backward_partials(C.byref(a), C.byref(n)))
    # the per-CTA partial rows of all outputs live in ONE [k, G, d] buffer, finished by ONE column reduce over axis 1
    k = 3 if want_dx_sum else 2
    parts = DeviceTensor.empty((k, n.value, d))
    pg, pb = parts.slice([(0, 1), (0, n.value), (0, d)]).reshape((n.value, d)), parts.slice([(1, 2), (0, n.value), (0, d)]).reshape((n.value, d))
    pgd, pbd = pg.desc(), pb.desc()
    gm = gamma.desc() if gamma is not None else None
    if not want_dx_sum:
        check(lib.b200_launch_layer_norm_backward(C.byref(a), C.byref(g), C.byref(gm) if gm is not None else None,
                                                  float(eps), C.byref(o), C.byref(pgd), C.byref(pbd), None))
    else:
        pdd = parts.slice([(2, 3), (0, n.value), (0, d)]).reshape((n.value, d)).desc()
        check(lib.b200_launch_layer_norm_backward_ex(C.byref(a), C.byref(g), C.byref(gm) if gm is not None else None,
                                                     float(eps), C.byref(o), C.byref(pgd), C.byref(pbd), C.byref(pdd), None))
    sums = float_sum_dim(parts, 1)                                  # [k, 1, d]
    outs = tuple(sums.slice([(i, i + 1), (0, 1), (0, d)]).reshape((d,)) for i in range(k))
    return (dx,) + outs


def attention(q: DeviceTensor, k: DeviceTensor, v: DeviceTensor, mask: DeviceTensor | None = None, scale: float | None = None,
              mask_value: float = float("-inf"), is_causal: bool = False, out: DeviceTensor | None = None,
              want_weights: bool = False):
    """ModuleOps::attention in one kernel (b200_launch_attention): softmax(q·kᵀ·scale [masked]) · v on
    [B,H,S,64] operands (strided head views are fine).  Returns `out`, or `(out, weights)`."""
    B, H, Sq, D = q.shape
    Sk, Dv = k.shape[2], v.shape[3]
    if scale is None:
        scale = 1.0 / float(np.sqrt(D))
    out = DeviceTensor.empty((B, H, Sq, Dv)) if out is None else out
    w = DeviceTensor.empty((B, H, Sq, Sk)) if want_weights else None
    qd, kd, vd, od = q.desc(), k.desc(), v.desc(), out.desc()
    md = mask.desc() if mask is not None else None
    wd = w.desc() if w is not None else None
    check(abi.load().b200_launch_attention(C.byref(qd), C.byref(kd), C.byref(vd), C.byref(md) if md is not None else None,
                                           float(scale), float(mask_value), 1 if is_causal else 0, C.byref(od),
                                           C.byref(wd) if wd is not None else None, None))
    return (out, w) if want_weights else out


def attention_backward(d_out: DeviceTensor, k: DeviceTensor, v: DeviceTensor, out: DeviceTensor, weights: DeviceTensor,
                       scale: float, is_causal: bool = False, dq: DeviceTensor | None = None):
    """(dq, ds) of the attention core from the saved weights and context (b200_launch_attention_backward);
    the caller finishes with dk = dsᵀ·q and dv = weightsᵀ·d_out."""
    B, H, Sq, D = d_out.shape
    dq = DeviceTensor.empty((B, H, Sq, D)) if dq is None else dq
    ds = DeviceTensor.empty(weights.shape)
    a, kd, vd, od, wd, qd, sd = (t.desc() for t in (d_out, k, v, out, weights, dq, ds))
    check(abi.load().b200_launch_attention_backward(C.byref(a), C.byref(kd), C.byref(vd), C.byref(od), C.byref(wd), float(scale),
                                                    1 if is_causal else 0, C.byref(qd), C.byref(sd), None))
    return dq, ds


def attention_flash(q: DeviceTensor, k: DeviceTensor, v: DeviceTensor, mask: DeviceTensor | None = None,
                    scale: float | None = None, mask_value: float = float("-inf"), is_causal: bool = False,
                    out: DeviceTensor | None = None):
    """ModuleOps::attention for training (b200_launch_attention_flash): returns (out, stats) where stats
    [B,H,Sq,4] holds the per-row softmax statistics the backward recomputes the weights from — no
    [B,H,Sq,Sk] tensor is written."""
    B, H, Sq, D = q.shape
    if scale is None:
        scale = 1.0 / float(np.sqrt(D))
    out = DeviceTensor.empty((B, H, Sq, v.shape[3])) if out is None else out
    stats = DeviceTensor.empty((B, H, Sq, 4))
    qd, kd, vd, od, sd = q.desc(), k.desc(), v.desc(), out.desc(), stats.desc()
    md = mask.desc() if mask is not None else None
    check(abi.load().b200_launch_attention_flash(C.byref(qd), C.byref(kd), C.byref(vd), C.byref(md) if md is not None else None,
                                                 float(scale), float(mask_value), 1 if is_causal else 0, C.byref(od),
                                                 C.byref(sd), None))
    return out, stats


def attention_flash_backward(d_out: DeviceTensor, q: DeviceTensor, k: DeviceTensor, v: DeviceTensor, out: DeviceTensor,
                             stats: DeviceTensor, mask: DeviceTensor | None, scale: float, mask_value: float,
                             is_causal: bool = False, dq=None, dk=None, dv=None):
    """(dq, dk, dv) of the attention core, weights recomputed from q, k and the saved statistics
    (b200_launch_attention_flash_backward)."""
    dq = DeviceTensor.empty(q.shape) if dq is None else dq
    dk = DeviceTensor.empty(k.shape) if dk is None else dk
    dv = DeviceTensor.empty(v.shape) if dv is None else dv
    ds = [t.desc() for t in (d_out, q, k, v, out, stats, dq, dk, dv)]
    md = mask.desc() if mask is not None else None
    check(abi.load().b200_launch_attention_flash_backward(
        C.byref(ds[0]), C.byref(ds[1]), C.byref(ds[2]), C.byref(ds[3]), C.byref(ds[4]), C.byref(ds[5]),
        C.byref(md) if md is not None else None, float(scale), float(mask_value), 1 if is_causal else 0,
        C.byref(ds[6]), C.byref(ds[7]), C.byref(ds[8]), None))
    return dq, dk, dv


def softmax_cross_entropy(logits: DeviceTensor, targets: DeviceTensor, grad_scale: float, inplace: bool = False):
    """(picked, dlogits): log_softmax(logits)[target] per row and (softmax - onehot) * grad_scale, one pass
    (b200_launch_softmax_cross_entropy).  `inplace` writes the gradient over the logits."""
    n, _ = logits.shape
    picked = DeviceTensor.empty((n,))
    dl = logits if inplace else DeviceTensor.empty(logits.shape)
    a, t, p, d = logits.desc(), targets.desc(), picked.desc(), dl.desc()
    check(abi.load().b200_launch_softmax_cross_entropy(C.byref(a), C.byref(t), float(grad_scale), C.byref(p), C.byref(d), None))
    return picked, dl
