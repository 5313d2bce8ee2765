"""TEST INFRASTRUCTURE ONLY — Python face of the CPU oracle.

Restates burn-ndarray (the reference's CPU backend, crates/burn-ndarray) for the
burn-b200 hot path on numpy arrays, delegating order-sensitive float arithmetic
to oracle/ndarray_oracle.c (compiled with `-O3 -march=native -ffp-contract=off`, no fast-math).
Function names follow the reference `FloatTensorOps` / `ActivationOps` /
`ModuleOps` entry points they restate (crates/burn-backend/src/backend/ops/).

Parity status: pinned — checked against the reference's golden vectors in
tests/test_oracle_golden.py (fixtures: tests/golden/burn_backend_tests.json).

Only tests/, bench.py's cpu_baseline / --impl reference legs and
__graft_entry__.smoke() may import this module.  The product never does.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SRC = HERE / "ndarray_oracle.c"
LIB = HERE / "_build" / "libndarray_oracle.so"

_lib = None

_BIN = {"add": 0, "sub": 1, "mul": 2, "div": 3, "rem": 4, "pow": 5, "min": 6, "max": 7, "remt": 8}
_UN = {n: i for i, n in enumerate(
    "exp log log1p sqrt tanh erf sin cos tan recip abs neg floor ceil round trunc sinh cosh asin "
    "acos atan asinh acosh atanh sign".split())}


def build(force: bool = False) -> Path:
    """Compiles the C restatement (building the checker is not using it)."""
    LIB.parent.mkdir(exist_ok=True)
    stamp = LIB.with_suffix(".host")
    host = _host_signature()
    fresh = LIB.exists() and LIB.stat().st_mtime > SRC.stat().st_mtime
    if fresh and not force and stamp.exists() and stamp.read_text() == host:
        return LIB
    # -O3 -march=native as BASELINE.md §3 states; contraction stays off (rustc never fuses a*b+c into an fma,
    # and the parity tests are bit-exact against these roundings).  -march=native binds the .so to this host's
    # ISA, hence the host stamp: a snapshot built elsewhere is rebuilt on the box that runs it.
    cmd = ["gcc", "-O3", "-ffp-contract=off", "-fno-fast-math", "-march=native", "-shared", "-fPIC",
           "-fvisibility=hidden", "-o", str(LIB), str(SRC), "-lm"]
    subprocess.run(cmd, check=True)
    stamp.write_text(host)
    return LIB


def _host_signature() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                import hashlib
                return hashlib.sha1(line.encode()).hexdigest()
    except OSError:
        pass
    return "unknown"


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(LIB))
        _lib.o_sum_f32.restype = C.c_float
    return _lib


def _f32(x) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(x, dtype=np.float32))


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


# ---------------------------------------------------------------- elementwise
def _binary(op: str, lhs, rhs) -> np.ndarray:
    lhs, rhs = np.asarray(lhs, dtype=np.float32), np.asarray(rhs, dtype=np.float32)
    shape = np.broadcast_shapes(lhs.shape, rhs.shape)
    a = _f32(np.broadcast_to(lhs, shape))
    b = _f32(np.broadcast_to(rhs, shape))
    out = np.empty(shape, dtype=np.float32)
    lib().o_binary_f32(_BIN[op], _p(a), _p(b), _p(out), C.c_size_t(out.size))
    return out


def _binary_scalar(op: str, lhs, s: float) -> np.ndarray:
    a = _f32(lhs)
    out = np.empty_like(a)
    lib().o_binary_scalar_f32(_BIN[op], _p(a), C.c_float(np.float32(s)), _p(out), C.c_size_t(out.size))
    return out


def _unary(op: str, x) -> np.ndarray:
    a = _f32(x)
    out = np.empty_like(a)
    lib().o_unary_f32(_UN[op], _p(a), _p(out), C.c_size_t(out.size))
    return out


# FloatTensorOps (crates/burn-backend/src/backend/ops/tensor.rs:183-306)
def float_add(a, b): return _binary("add", a, b)
def float_sub(a, b): return _binary("sub", a, b)
def float_mul(a, b): return _binary("mul", a, b)
def float_div(a, b): return _binary("div", a, b)
def float_remainder(a, b): return _binary("remt", a, b)    # base.rs:909-922 (the scalar form is a different formula)
def float_powf(a, b): return _binary("pow", a, b)
def float_add_scalar(a, s): return _binary_scalar("add", a, s)
def float_sub_scalar(a, s): return _binary_scalar("sub", a, s)
def float_mul_scalar(a, s): return _binary_scalar("mul", a, s)
def float_div_scalar(a, s): return _binary_scalar("div", a, s)
def float_remainder_scalar(a, s): return _binary_scalar("rem", a, s)
def float_powf_scalar(a, s): return _binary_scalar("pow", a, s)
def float_exp(a): return _unary("exp", a)
def float_log(a): return _unary("log", a)
def float_log1p(a): return _unary("log1p", a)
def float_sqrt(a): return _unary("sqrt", a)
def float_tanh(a): return _unary("tanh", a)
def float_erf(a): return _unary("erf", a)
def float_sin(a): return _unary("sin", a)
def float_cos(a): return _unary("cos", a)
def float_tan(a): return _unary("tan", a)
def float_recip(a): return _unary("recip", a)
def float_abs(a): return _unary("abs", a)
def float_neg(a): return _unary("neg", a)
def float_floor(a): return _unary("floor", a)
def float_ceil(a): return _unary("ceil", a)
def float_round(a): return _unary("round", a)
def float_trunc(a): return _unary("trunc", a)
def float_sign(a): return _unary("sign", a)


def float_clamp(a, lo, hi):
    a = _f32(a)
    return np.where(np.isnan(a), a, np.minimum(np.maximum(a, np.float32(lo)), np.float32(hi))).astype(np.float32)


# comparisons → bool (crates/burn-backend/src/backend/ops/tensor.rs:675-850)
def float_equal(a, b): return np.equal(_f32(a), np.asarray(b, dtype=np.float32))
def float_not_equal(a, b): return np.not_equal(_f32(a), np.asarray(b, dtype=np.float32))
def float_greater(a, b): return np.greater(_f32(a), np.asarray(b, dtype=np.float32))
def float_greater_equal(a, b): return np.greater_equal(_f32(a), np.asarray(b, dtype=np.float32))
def float_lower(a, b): return np.less(_f32(a), np.asarray(b, dtype=np.float32))
def float_lower_equal(a, b): return np.less_equal(_f32(a), np.asarray(b, dtype=np.float32))


# mask ops (crates/burn-ndarray/src/ops/base.rs:78-104): mask broadcast to the tensor
def float_mask_fill(a, mask, value: float) -> np.ndarray:
    a = np.asarray(a, dtype=np.float32)
    mask = np.asarray(mask).astype(bool)
    shape = np.broadcast_shapes(a.shape, mask.shape)
    return np.where(np.broadcast_to(mask, shape), np.float32(value), np.broadcast_to(a, shape)).astype(np.float32)


def float_mask_where(a, mask, source) -> np.ndarray:
    a = np.asarray(a, dtype=np.float32)
    mask = np.asarray(mask).astype(bool)
    source = np.asarray(source, dtype=np.float32)
    shape = np.broadcast_shapes(a.shape, mask.shape, source.shape)
    return np.where(np.broadcast_to(mask, shape), np.broadcast_to(source, shape),
                    np.broadcast_to(a, shape)).astype(np.float32)


# ---------------------------------------------------------------- reductions
def _split(shape, dim):
    outer = int(np.prod(shape[:dim], dtype=np.int64))
    inner = int(np.prod(shape[dim + 1:], dtype=np.int64))
    return outer, int(shape[dim]), inner


def _keepdim(shape, dim):
    s = list(shape)
    s[dim] = 1
    return tuple(s)


def float_sum(a) -> np.ndarray:
    """Shape [1] (crates/burn-ndarray/src/ops/base.rs:940-943)."""
    a = _f32(a)
    return np.array([lib().o_sum_f32(_p(a), C.c_size_t(a.size))], dtype=np.float32)


def float_mean(a) -> np.ndarray:
    a = _f32(a)
    return (float_sum(a) / np.float32(a.size)).astype(np.float32)


def _axis(fn_name: str, a, dim: int) -> np.ndarray:
    a = _f32(a)
    if dim < 0 or dim >= a.ndim:
        raise IndexError(f"dim {dim} out of range for rank {a.ndim}")
    outer, R, inner = _split(a.shape, dim)
    out = np.empty(_keepdim(a.shape, dim), dtype=np.float32)
    getattr(lib(), fn_name)(_p(a), C.c_size_t(outer), C.c_size_t(R), C.c_size_t(inner), _p(out))
    return out


def float_sum_dim(a, dim): return _axis("o_sum_axis_f32", a, dim)
def float_mean_dim(a, dim): return _axis("o_mean_axis_f32", a, dim)
def float_prod_dim(a, dim): return _axis("o_prod_axis_f32", a, dim)


def _arg(a, dim: int, is_min: bool) -> np.ndarray:
    a = np.ascontiguousarray(a)
    if dim < 0 or dim >= a.ndim:
        raise IndexError(f"dim {dim} out of range for rank {a.ndim}")
    if a.shape[dim] == 0:
        raise ValueError("Cannot compute arg over an empty axis")
    outer, R, inner = _split(a.shape, dim)
    out = np.empty(_keepdim(a.shape, dim), dtype=np.int64)
    if a.dtype.kind == "f":
        a = _f32(a)
        lib().o_arg_f32(_p(a), C.c_size_t(outer), C.c_size_t(R), C.c_size_t(inner), int(is_min), _p(out))
    else:
        a = np.ascontiguousarray(a, dtype=np.int64)
        lib().o_arg_i64(_p(a), C.c_size_t(outer), C.c_size_t(R), C.c_size_t(inner), int(is_min), _p(out))
    return out


def float_argmax(a, dim): return _arg(a, dim, False)
def float_argmin(a, dim): return _arg(a, dim, True)


def float_max_dim(a, dim):
    """default = gather(dim, x, argmax(x)) (crates/burn-backend/src/backend/ops/tensor.rs:1609-1614)."""
    return float_gather(dim, a, float_argmax(a, dim))


def float_min_dim(a, dim):
    return float_gather(dim, a, float_argmin(a, dim))


# ---------------------------------------------------------------- matmul
def float_matmul(lhs, rhs) -> np.ndarray:
    """Broadcast-batched GEMM (crates/burn-ndarray/src/ops/matmul.rs:9-183)."""
    lhs, rhs = np.asarray(lhs, dtype=np.float32), np.asarray(rhs, dtype=np.float32)
    if lhs.ndim != rhs.ndim or lhs.ndim < 2:
        raise ValueError("matmul operands must have the same rank >= 2")
    M, K = lhs.shape[-2:]
    K2, N = rhs.shape[-2:]
    if K != K2:
        raise ValueError(f"matmul inner dims differ: {K} vs {K2}")
    for x, y in zip(lhs.shape[:-2], rhs.shape[:-2]):
        if x != y and x != 1 and y != 1:
            raise ValueError("matmul batch dims are not broadcastable")
    batch = np.broadcast_shapes(lhs.shape[:-2], rhs.shape[:-2])
    lb = np.broadcast_to(lhs, batch + (M, K))
    rb = np.broadcast_to(rhs, batch + (K, N))
    out = np.empty(batch + (M, N), dtype=np.float32)
    for idx in np.ndindex(*batch):
        a = _f32(lb[idx])
        b = _f32(rb[idx])
        c = np.empty((M, N), dtype=np.float32)
        lib().o_sgemm(_p(a), C.c_int64(K), C.c_int64(1), _p(b), C.c_int64(N), C.c_int64(1), _p(c),
                      C.c_int64(M), C.c_int64(N), C.c_int64(K))
        out[idx] = c
    return out


# ---------------------------------------------------------------- indexing
def _idx_layout(shape, dim, idx_len):
    outer, D, inner = _split(shape, dim)
    return outer, D, idx_len, inner


def float_gather(dim, t, indices) -> np.ndarray:
    t = _f32(t)
    idx = np.ascontiguousarray(indices, dtype=np.int64)
    outer, D, inner = _split(t.shape, dim)
    if idx.shape[:dim] != t.shape[:dim] or idx.shape[dim + 1:] != t.shape[dim + 1:]:
        raise ValueError("gather: indices must match tensor on all dims but `dim`")
    out = np.empty(idx.shape, dtype=np.float32)
    lib().o_gather_f32(_p(t), _p(idx), _p(out), C.c_size_t(outer), C.c_size_t(D),
                       C.c_size_t(idx.shape[dim]), C.c_size_t(inner))
    return out


def float_scatter_add(dim, t, indices, value) -> np.ndarray:
    t = _f32(t).copy()
    idx = np.ascontiguousarray(indices, dtype=np.int64)
    v = _f32(value)
    if idx.shape != v.shape:
        raise ValueError("scatter: indices and value shapes differ")
    outer, D, inner = _split(t.shape, dim)
    lib().o_scatter_add_f32(_p(t), _p(idx), _p(v), C.c_size_t(outer), C.c_size_t(D),
                            C.c_size_t(idx.shape[dim]), C.c_size_t(inner))
    return t


def float_select(t, dim, indices) -> np.ndarray:
    t = _f32(t)
    idx = np.ascontiguousarray(indices, dtype=np.int64)
    outer, D, inner = _split(t.shape, dim)
    shape = list(t.shape)
    shape[dim] = idx.size
    out = np.empty(shape, dtype=np.float32)
    lib().o_select_f32(_p(t), _p(idx), _p(out), C.c_size_t(outer), C.c_size_t(D), C.c_size_t(idx.size),
                       C.c_size_t(inner))
    return out


def float_select_add(t, dim, indices, value) -> np.ndarray:
    t = _f32(t).copy()
    idx = np.ascontiguousarray(indices, dtype=np.int64)
    v = _f32(value)
    outer, D, inner = _split(t.shape, dim)
    lib().o_select_add_f32(_p(t), _p(idx), _p(v), C.c_size_t(outer), C.c_size_t(D), C.c_size_t(idx.size),
                           C.c_size_t(inner))
    return t


# ---------------------------------------------------------------- data movement (unit steps; pure copies, so numpy
# slicing IS the restatement — there is no arithmetic whose order could differ)
def _ranges(shape, ranges):
    """Tensor::slice canonicalisation (crates/burn-tensor/src/tensor/api/base.rs `slice`; burn-std Slice::to_range):
    negative bounds count from the end, ends clamp to the dimension, unspecified trailing dims are full."""
    out = []
    for i, n in enumerate(shape):
        if i < len(ranges):
            lo, hi = ranges[i]
            lo = lo + n if lo < 0 else lo
            hi = n if hi is None else (hi + n if hi < 0 else hi)
            lo, hi = min(max(lo, 0), n), min(max(hi, 0), n)
            out.append((lo, max(hi, lo)))
        else:
            out.append((0, n))
    return out


def float_slice(t, ranges) -> np.ndarray:
    """NdArrayOps::slice (crates/burn-ndarray/src/ops/base.rs:62-65)."""
    t = _f32(t)
    return np.ascontiguousarray(t[tuple(slice(a, b) for a, b in _ranges(t.shape, ranges))])


def float_slice_assign(t, ranges, value) -> np.ndarray:
    """NdArrayOps::slice_assign (crates/burn-ndarray/src/ops/base.rs:67-76): owned copy, slice_mut().assign(value)."""
    out = _f32(t).copy()
    out[tuple(slice(a, b) for a, b in _ranges(out.shape, ranges))] = _f32(value)
    return out


def float_cat(tensors, dim: int) -> np.ndarray:
    """NdArrayOps::cat (crates/burn-ndarray/src/ops/base.rs:437-440)."""
    return np.concatenate([_f32(t) for t in tensors], axis=dim)


def float_flip(t, axes) -> np.ndarray:
    """NdArrayOps::flip (crates/burn-ndarray/src/ops/base.rs:507-529): step -1 on the named axes, then owned."""
    return np.ascontiguousarray(np.flip(_f32(t), tuple(axes)))


def float_repeat_dim(t, dim: int, times: int) -> np.ndarray:
    """float_repeat_dim default (crates/burn-backend/src/backend/ops/tensor.rs:161-163 → repeat_with_slice_assign):
    `times` copies laid one after the other along `dim`."""
    return np.concatenate([_f32(t)] * int(times), axis=dim)


# ---------------------------------------------------------------- composites
SQRT_2 = np.float32(1.4142135623730951)


def relu(x):
    """crates/burn-backend/src/backend/ops/activation.rs:37-42"""
    return float_mask_fill(x, float_lower_equal(x, 0.0), 0.0)


def gelu(x):
    """crates/burn-backend/src/backend/ops/activation.rs:69-76"""
    t = float_div_scalar(x, SQRT_2)
    t = float_erf(t)
    t = float_add_scalar(t, 1.0)
    t = float_mul(x, t)
    return float_div_scalar(t, 2.0)


def sigmoid(x):
    """1 / (1 + exp(-x)) evaluated in f32 like the reference default."""
    return float_recip(float_add_scalar(float_exp(float_neg(x)), 1.0))


def softmax(x, dim):
    """crates/burn-backend/src/backend/ops/activation.rs:250-256"""
    m = float_max_dim(x, dim)
    e = float_exp(float_sub(x, m))
    return float_div(e, float_sum_dim(e, dim))


def log_softmax(x, dim):
    """crates/burn-backend/src/backend/ops/activation.rs:271-276"""
    m = float_max_dim(x, dim)
    shifted = float_sub(x, m)
    lse = float_log(float_sum_dim(float_exp(shifted), dim))
    return float_sub(shifted, lse)


def layer_norm(x, gamma, beta, eps: float):
    """crates/burn-backend/src/backend/ops/modules/base.rs:846-877"""
    x = _f32(x)
    last = x.ndim - 1
    mean = float_mean_dim(x, last)
    centered = float_sub(x, mean)
    var = float_mean_dim(float_mul(centered, centered), last)
    denom = float_sqrt(float_add_scalar(var, np.float32(eps)))
    y = float_div(centered, denom)
    if gamma is not None:
        y = float_mul(y, np.asarray(gamma, dtype=np.float32).reshape((1,) * last + (-1,)))
    if beta is not None:
        y = float_add(y, np.asarray(beta, dtype=np.float32).reshape((1,) * last + (-1,)))
    return y


def linear(x, weight, bias=None):
    """crates/burn-backend/src/backend/ops/modules/linear.rs:17-64: x·W (+ b), W is [d_in, d_out]."""
    x = _f32(x)
    w = np.asarray(weight, dtype=np.float32)
    y = float_matmul(x.reshape((-1, x.shape[-1]))[None], w[None])[0].reshape(x.shape[:-1] + (w.shape[-1],))
    if bias is not None:
        y = float_add(y, np.asarray(bias, dtype=np.float32).reshape((1,) * (y.ndim - 1) + (-1,)))
    return y


# ---------------------------------------------------------------- tolerance
def approx_eq_mask(x, y, rel: float, abs_: float) -> np.ndarray:
    """burn_std Tolerance semantics: |x-y| < max(rel*|x+y|, abs) (crates/burn-std/src/data/compare.rs:10-27);
    equal values (incl. same-signed inf) and NaN==NaN pass."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    both_nan = np.isnan(x) & np.isnan(y)
    eq = x == y
    with np.errstate(invalid="ignore"):
        close = np.abs(x - y) < np.maximum(rel * np.abs(x + y), abs_)
    return both_nan | eq | close


def bench_chain_unfused(a, b, c, m):
    a, b, c = _f32(a), _f32(b), _f32(c)
    m = np.ascontiguousarray(m, dtype=np.uint8)
    out = np.empty_like(a)
    tmp = np.empty(2 * a.size, dtype=np.float32)
    lib().o_bench_chain_unfused(_p(a), _p(b), _p(c), _p(m), _p(out), _p(tmp), C.c_size_t(a.size))
    return out


def bench_step_fused(a, b, c, m, threads: int = 1):
    """o_bench_step_fused over row blocks, one block per thread (ctypes releases the GIL).  Returns
    (y, row_sum, row_mean, row_argmax, col_sum, total).  A hand-fused CPU baseline, not burn-ndarray's execution."""
    from concurrent.futures import ThreadPoolExecutor
    a, b, c = _f32(a), _f32(b), _f32(c)
    m = np.ascontiguousarray(m, dtype=np.uint8)
    rows, cols = a.shape
    y = np.empty_like(a)
    rs, rm = np.empty(rows, np.float32), np.empty(rows, np.float32)
    am = np.empty(rows, np.int64)
    threads = max(1, min(threads, rows))
    cs = np.zeros((threads, cols), np.float32)
    tot = np.zeros(threads, np.float64)
    edges = np.linspace(0, rows, threads + 1).astype(int)
    fn = lib().o_bench_step_fused

    def work(t):
        lo, hi = int(edges[t]), int(edges[t + 1])
        if hi > lo:
            fn(_p(a[lo:hi]), _p(b[lo:hi]), _p(c[lo:hi]), _p(m[lo:hi]), _p(y[lo:hi]), _p(rs[lo:hi]), _p(rm[lo:hi]),
               _p(am[lo:hi]), _p(cs[t]), _p(tot[t:t + 1]), C.c_size_t(hi - lo), C.c_size_t(cols))

    if threads == 1:
        work(0)
    else:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(work, range(threads)))
    return y, rs, rm, am, cs.sum(axis=0, dtype=np.float32), float(tot.sum())
