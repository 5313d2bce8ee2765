"""Evaluates the expression-tree goldens (tests/golden/burn_backend_tests_expr.json, extracted from the reference's
own tests by scripts/extract_goldens.py) on a backend: the CPU oracle (`-m "not gpu"`, pins the oracle) or the
CUDA library through the C ABI (`-m gpu`).

A tree node is {"op": "lit", "kind": float|int|bool, "value": nested list} or {"op": name, "x": node, "args": [...]}
with `name` a public Tensor-API method of the reference (crates/burn-tensor/src/tensor/api/*.rs); both backends map
it onto their FloatTensorOps-level functions exactly as the reference's Tensor API does (negative dims are
canonicalised, `max_dim_with_indices` = (max_dim, argmax), `scatter(.., Add)` = scatter_add ...).
"""
from __future__ import annotations

import json
import math
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden" / "burn_backend_tests_expr.json"


def load_cases():
    if not GOLDEN.exists():
        return []
    return json.loads(GOLDEN.read_text())["cases"]


# ------------------------------------------------------------------ literal encoding (JSON has no nan / inf)
def encode_lit(x):
    if isinstance(x, list):
        return [encode_lit(e) for e in x]
    if isinstance(x, float):
        if math.isnan(x):
            return "nan"
        if math.isinf(x):
            return "inf" if x > 0 else "-inf"
    return x


def decode_lit(x):
    if isinstance(x, list):
        return [decode_lit(e) for e in x]
    if isinstance(x, str):
        return {"nan": math.nan, "inf": math.inf, "-inf": -math.inf}[x]
    return x


def encode_tree(n):
    if isinstance(n, dict):
        out = {}
        for k, v in n.items():
            out[k] = encode_lit(v) if k == "value" else encode_tree(v)
        return out
    if isinstance(n, list):
        return [encode_tree(e) for e in n]
    if isinstance(n, tuple):
        raise NotImplementedError(f"argument {n!r}")
    return encode_lit(n)


NP_KIND = {"float": np.float32, "int": np.int64, "bool": np.bool_}


def check(case, got):
    want = np.array(decode_lit(case["expected"]), dtype=np.float64)
    got = np.asarray(got)
    if got.shape != want.shape and got.size == want.size and (got.ndim == 0 or want.ndim == 0 or got.size == 1):
        got = got.reshape(want.shape)          # scalar results: the reference yields shape [1]
    assert got.shape == want.shape, f"{case['name']} ({case['cite']}): shape {got.shape} != {want.shape}"
    if case["tol"] == "exact":
        # TensorData::assert_eq(strict=false) converts the expected literals to the tensor's dtype
        w = want.astype(got.dtype) if got.dtype.kind == "f" else want
        g = got.astype(np.float64)
        same = (got == w) | (np.isnan(g) & np.isnan(want))
        assert same.all(), f"{case['name']} ({case['cite']}): got {got.tolist()} want {want.tolist()}"
    else:
        from oracle import oracle
        rel, abs_ = case["tol"]
        ok = oracle.approx_eq_mask(got.astype(np.float64), want, rel, abs_)
        assert ok.all(), f"{case['name']} ({case['cite']}): got {got.tolist()} want {want.tolist()}"


# ------------------------------------------------------------------ evaluation
def evaluate(node, be):
    """Returns a numpy array.  `be` is a backend (OracleBackend / DeviceBackend)."""
    return be.to_numpy(_eval(node, be))


def _eval(node, be):
    op = node["op"]
    if op == "lit":
        return be.literal(np.array(decode_lit(node["value"]), dtype=NP_KIND[node["kind"]]), node["kind"])
    x = _eval(node["x"], be)
    args = [_arg(a, be) for a in node.get("args", [])]
    fn = getattr(be, "op_" + op, None)
    if fn is None:
        raise NotImplementedError(op)
    if "out" in node:
        return fn(x, *args, out=node["out"])
    return fn(x, *args)


def _arg(a, be):
    if isinstance(a, dict):
        if "op" in a:
            return _eval(a, be)
        if "range" in a:
            return tuple(a["range"])
    if isinstance(a, list) and a and all(isinstance(e, dict) for e in a):
        return [_arg(e, be) for e in a]          # a list of tensors (cat) or of ranges (slice)
    return decode_lit(a)


def _dim(d, rank):
    d = int(d)
    return d + rank if d < 0 else d


def _rng(ranges):
    if isinstance(ranges, tuple):
        ranges = [ranges]
    out = []
    for r in ranges:
        if not (isinstance(r, tuple) and len(r) == 2):
            raise NotImplementedError(f"slice argument {r!r}")
        out.append((int(r[0]), int(r[1])))
    return out


class _Common:
    """Tensor-API level compositions shared by both backends (crates/burn-tensor/src/tensor/api/numeric.rs etc.)."""

    # the ops below are expressed through the backend's primitive table `P`
    def _k(self, x):
        raise NotImplementedError

    def op_max_dim_with_indices(self, x, d, out):
        return self.op_max_dim(x, d) if out == 0 else self.op_argmax(x, d)

    def op_min_dim_with_indices(self, x, d, out):
        return self.op_min_dim(x, d) if out == 0 else self.op_argmin(x, d)

    def op_square(self, x):
        return self.op_mul(x, x)

    def op_flatten(self, x, a, b):
        a, b = _dim(a, x.ndim), _dim(b, x.ndim)
        shape = list(x.shape)
        return self.op_reshape(x, shape[:a] + [int(np.prod(shape[a:b + 1]))] + shape[b + 1:])

    def op_unsqueeze(self, x):
        raise NotImplementedError("unsqueeze::<D2>() needs the generic rank")

    def op_narrow(self, x, dim, start, length):
        dim = _dim(dim, x.ndim)
        return self.op_slice(x, [(0, n) for n in x.shape[:dim]] + [(int(start), int(start) + int(length))])

    def op_slice_fill(self, x, ranges, v):
        raise NotImplementedError("slice_fill")


# ------------------------------------------------------------------ oracle backend (numpy arrays)
class OracleBackend(_Common):
    def __init__(self):
        from oracle import oracle
        self.o = oracle

    def literal(self, a, kind):
        return a

    def to_numpy(self, t):
        return np.asarray(t)

    def _f(self, x):
        if x.dtype != np.float32:
            raise NotImplementedError(f"{x.dtype} tensor op")
        return np.ascontiguousarray(x)

    # views / casts
    def op_reshape(self, x, shape):
        shape = [int(s) for s in shape]
        if -1 in shape or 0 in shape:
            known = int(np.prod([s for s in shape if s > 0]))
            shape = [x.shape[i] if s == 0 else (x.size // known if s == -1 else s) for i, s in enumerate(shape)]
        return np.ascontiguousarray(x).reshape(shape)

    def op_transpose(self, x):
        return np.swapaxes(x, -1, -2)

    def op_swap_dims(self, x, a, b):
        return np.swapaxes(x, _dim(a, x.ndim), _dim(b, x.ndim))

    def op_permute(self, x, axes):
        return np.transpose(x, [_dim(a, x.ndim) for a in axes])

    def op_float(self, x):
        return x.astype(np.float32)

    def op_int(self, x):
        if x.dtype == np.float32:
            return np.trunc(x).astype(np.int64)
        return x.astype(np.int64)

    def op_expand(self, x, shape):
        shape = [int(s) for s in shape]
        lead = len(shape) - x.ndim
        full = [x.shape[i - lead] if s == -1 else s for i, s in enumerate(shape)]
        return np.broadcast_to(x, full)

    def op_flip(self, x, axes): return self.o.float_flip(self._f(x), [_dim(a, x.ndim) for a in axes])

    def op_repeat_dim(self, x, dim, times): return self.o.float_repeat_dim(self._f(x), _dim(dim, x.ndim), int(times))
    def op_slice(self, x, ranges): return self.o.float_slice(self._f(x), _rng(ranges))
    def op_slice_assign(self, x, ranges, v): return self.o.float_slice_assign(self._f(x), _rng(ranges), self._f(v))
    def op_cat(self, x, rest, dim): return self.o.float_cat([self._f(t) for t in [x] + list(rest)], _dim(dim, x.ndim))

    # elementwise
    def _bin(self, name, x, y):
        return getattr(self.o, "float_" + name)(self._f(x), self._f(y))

    def op_add(self, x, y): return self._bin("add", x, y)
    def op_sub(self, x, y): return self._bin("sub", x, y)
    def op_mul(self, x, y): return self._bin("mul", x, y)
    def op_div(self, x, y): return self._bin("div", x, y)
    def op_remainder(self, x, y): return self._bin("remainder", x, y)
    def op_powf(self, x, y): return self._bin("powf", x, y)
    def op_add_scalar(self, x, s): return self.o.float_add_scalar(self._f(x), s)
    def op_sub_scalar(self, x, s): return self.o.float_sub_scalar(self._f(x), s)
    def op_mul_scalar(self, x, s): return self.o.float_mul_scalar(self._f(x), s)
    def op_div_scalar(self, x, s): return self.o.float_div_scalar(self._f(x), s)
    def op_remainder_scalar(self, x, s): return self.o.float_remainder_scalar(self._f(x), s)
    def op_powf_scalar(self, x, s): return self.o.float_powf_scalar(self._f(x), s)

    def op_powi_scalar(self, x, s):
        # float_powi_scalar (crates/burn-backend/src/backend/ops/tensor.rs:1095-1104)
        s = int(s)
        x = self._f(x)
        if s == 0:
            return np.ones_like(x)
        if s == 1:
            return x
        if s == 2:
            return self.o.float_mul(x, x)
        if s == -1:
            return self.o.float_recip(x)
        if s == -2:
            return self.o.float_recip(self.o.float_mul(x, x))
        return self.o.float_powf_scalar(x, float(s))

    def _un(name):   # noqa: N805
        def f(self, x):
            return getattr(self.o, "float_" + name)(self._f(x))
        return f

    for _n in ("exp", "log", "log1p", "sqrt", "abs", "neg", "recip", "tanh", "erf", "sin", "cos", "tan", "floor", "ceil", "round",
               "trunc", "sign"):
        locals()["op_" + _n] = _un(_n)
    del _n, _un

    def op_clamp(self, x, lo, hi): return self.o.float_clamp(self._f(x), lo, hi)
    def op_clamp_min(self, x, lo): return np.where(self._f(x) < np.float32(lo), np.float32(lo), x).astype(np.float32)
    def op_clamp_max(self, x, hi): return np.where(self._f(x) > np.float32(hi), np.float32(hi), x).astype(np.float32)
    def op_is_nan(self, x): return np.isnan(self._f(x))
    def op_is_inf(self, x): return np.isinf(self._f(x))

    # comparisons
    def op_equal(self, x, y): return self.o.float_equal(self._f(x), self._f(y))
    def op_not_equal(self, x, y): return self.o.float_not_equal(self._f(x), self._f(y))
    def op_greater(self, x, y): return self.o.float_greater(self._f(x), self._f(y))
    def op_greater_equal(self, x, y): return self.o.float_greater_equal(self._f(x), self._f(y))
    def op_lower(self, x, y): return self.o.float_lower(self._f(x), self._f(y))
    def op_lower_equal(self, x, y): return self.o.float_lower_equal(self._f(x), self._f(y))
    def op_equal_elem(self, x, s): return self.o.float_equal(self._f(x), np.float32(s))
    def op_not_equal_elem(self, x, s): return self.o.float_not_equal(self._f(x), np.float32(s))
    def op_greater_elem(self, x, s): return self.o.float_greater(self._f(x), np.float32(s))
    def op_greater_equal_elem(self, x, s): return self.o.float_greater_equal(self._f(x), np.float32(s))
    def op_lower_elem(self, x, s): return self.o.float_lower(self._f(x), np.float32(s))
    def op_lower_equal_elem(self, x, s): return self.o.float_lower_equal(self._f(x), np.float32(s))
    def op_mask_fill(self, x, m, v): return self.o.float_mask_fill(self._f(x), np.asarray(m, bool), v)
    def op_mask_where(self, x, m, src): return self.o.float_mask_where(self._f(x), np.asarray(m, bool), self._f(src))

    # reductions
    def op_sum(self, x): return self.o.float_sum(self._f(x))
    def op_mean(self, x): return self.o.float_mean(self._f(x))
    def op_sum_dim(self, x, d): return self.o.float_sum_dim(self._f(x), _dim(d, x.ndim))
    def op_mean_dim(self, x, d): return self.o.float_mean_dim(self._f(x), _dim(d, x.ndim))
    def op_prod_dim(self, x, d): return self.o.float_prod_dim(self._f(x), _dim(d, x.ndim))
    def op_max_dim(self, x, d): return self.o.float_max_dim(self._f(x), _dim(d, x.ndim))
    def op_min_dim(self, x, d): return self.o.float_min_dim(self._f(x), _dim(d, x.ndim))
    def op_argmax(self, x, d): return self.o.float_argmax(self._f(x), _dim(d, x.ndim))
    def op_argmin(self, x, d): return self.o.float_argmin(self._f(x), _dim(d, x.ndim))

    def op_max(self, x):
        # float_max default: argmax over the flattened tensor + gather (crates/burn-backend/src/backend/ops/tensor.rs:1592-1602)
        f = self._f(x).reshape(-1)
        return self.o.float_max_dim(f, 0)

    def op_min(self, x):
        f = self._f(x).reshape(-1)
        return self.o.float_min_dim(f, 0)

    # contraction / indexing
    def op_matmul(self, x, y):
        x, y = self._f(x), self._f(y)
        if x.ndim == 1 or y.ndim == 1:
            raise NotImplementedError("1-D matmul")
        return self.o.float_matmul(x, y)

    def op_gather(self, x, d, idx): return self.o.float_gather(_dim(d, x.ndim), self._f(x), np.asarray(idx, np.int64))
    def op_scatter_add(self, x, d, idx, v): return self.o.float_scatter_add(_dim(d, x.ndim), self._f(x), np.asarray(idx, np.int64), self._f(v))
    def op_select(self, x, d, idx): return self.o.float_select(self._f(x), _dim(d, x.ndim), np.asarray(idx, np.int64))
    def op_select_add(self, x, d, idx, v): return self.o.float_select_add(self._f(x), _dim(d, x.ndim), np.asarray(idx, np.int64), self._f(v))

    # activations
    def op_relu(self, x): return self.o.relu(self._f(x))
    def op_gelu(self, x): return self.o.gelu(self._f(x))
    def op_sigmoid(self, x): return self.o.sigmoid(self._f(x))
    def op_softmax(self, x, d): return self.o.softmax(self._f(x), _dim(d, x.ndim))
    def op_log_softmax(self, x, d): return self.o.log_softmax(self._f(x), _dim(d, x.ndim))


# ------------------------------------------------------------------ device backend (C ABI)
class DeviceBackend(_Common):
    def __init__(self):
        from burn_b200 import _abi as abi
        from burn_b200 import ops
        from burn_b200.device import DeviceTensor
        self.abi, self.ops, self.DT = abi, ops, DeviceTensor

    def literal(self, a, kind):
        return self.DT.from_numpy(a)

    def to_numpy(self, t):
        return t.numpy()

    def _f(self, x):
        if x.dtype != self.abi.F32:
            raise NotImplementedError(f"dtype {x.dtype} tensor op")
        return x

    def op_reshape(self, x, shape):
        shape = [int(s) for s in shape]
        if -1 in shape or 0 in shape:
            known = int(np.prod([s for s in shape if s > 0]))
            shape = [x.shape[i] if s == 0 else (x.numel // known if s == -1 else s) for i, s in enumerate(shape)]
        return x.reshape(shape)

    def op_transpose(self, x): return x.swap_dims(x.ndim - 1, x.ndim - 2)
    def op_swap_dims(self, x, a, b): return x.swap_dims(_dim(a, x.ndim), _dim(b, x.ndim))
    def op_permute(self, x, axes): return x.permute([_dim(a, x.ndim) for a in axes])

    def op_expand(self, x, shape):
        shape = [int(s) for s in shape]
        lead = len(shape) - x.ndim
        full = [x.shape[i - lead] if s == -1 else s for i, s in enumerate(shape)]
        return x.reshape((1,) * lead + tuple(x.shape)).expand(full)

    def op_float(self, x): return self.ops.int_into_float(x) if x.dtype != self.abi.F32 else x
    def op_int(self, x): return self.ops.float_into_int(x, self.abi.I64) if x.dtype == self.abi.F32 else x
    def op_flip(self, x, axes): return self.ops.float_flip(self._f(x), [_dim(a, x.ndim) for a in axes])
    def op_repeat_dim(self, x, dim, times): return self.ops.float_repeat_dim(self._f(x), _dim(dim, x.ndim), int(times))
    def op_slice(self, x, ranges): return self.ops.float_slice(self._f(x), _rng(ranges))
    def op_slice_assign(self, x, ranges, v): return self.ops.float_slice_assign(self._f(x), _rng(ranges), self._f(v))
    def op_cat(self, x, rest, dim): return self.ops.float_cat([self._f(t) for t in [x] + list(rest)], _dim(dim, x.ndim))

    def _bin(name):   # noqa: N805
        def f(self, x, y):
            return getattr(self.ops, "float_" + name)(self._f(x), self._f(y))
        return f

    for _n in ("add", "sub", "mul", "div", "remainder", "powf", "equal", "not_equal", "greater", "greater_equal", "lower", "lower_equal"):
        locals()["op_" + _n] = _bin(_n)
    del _n, _bin

    def _sc(name):   # noqa: N805
        def f(self, x, s):
            return getattr(self.ops, "float_" + name)(self._f(x), s)
        return f

    for _n in ("add_scalar", "sub_scalar", "mul_scalar", "div_scalar", "remainder_scalar", "powf_scalar", "powi_scalar", "equal_elem",
               "not_equal_elem", "greater_elem", "greater_equal_elem", "lower_elem", "lower_equal_elem", "clamp_min", "clamp_max"):
        locals()["op_" + _n] = _sc(_n)
    del _n, _sc

    def _un(name):   # noqa: N805
        def f(self, x):
            return getattr(self.ops, "float_" + name)(self._f(x))
        return f

    for _n in ("exp", "log", "log1p", "sqrt", "abs", "neg", "recip", "tanh", "erf", "sin", "cos", "tan", "floor", "ceil", "round",
               "trunc", "sign", "is_nan", "is_inf", "sum", "mean", "max", "min"):
        locals()["op_" + _n] = _un(_n)
    del _n, _un

    def op_clamp(self, x, lo, hi): return self.ops.float_clamp(self._f(x), lo, hi)
    def op_mask_fill(self, x, m, v): return self.ops.float_mask_fill(self._f(x), m, v)
    def op_mask_where(self, x, m, src): return self.ops.float_mask_where(self._f(x), m, self._f(src))

    def _red(name):   # noqa: N805
        def f(self, x, d):
            return getattr(self.ops, "float_" + name)(self._f(x), _dim(d, x.ndim))
        return f

    for _n in ("sum_dim", "mean_dim", "prod_dim", "max_dim", "min_dim", "argmax", "argmin"):
        locals()["op_" + _n] = _red(_n)
    del _n, _red

    def op_matmul(self, x, y):
        if x.ndim == 1 or y.ndim == 1:
            raise NotImplementedError("1-D matmul")
        return self.ops.float_matmul(self._f(x), self._f(y))

    def op_gather(self, x, d, idx): return self.ops.float_gather(_dim(d, x.ndim), self._f(x), idx)
    def op_scatter_add(self, x, d, idx, v): return self.ops.float_scatter_add(_dim(d, x.ndim), self._f(x), idx, self._f(v))
    def op_select(self, x, d, idx): return self.ops.float_select(self._f(x), _dim(d, x.ndim), idx)
    def op_select_add(self, x, d, idx, v): return self.ops.float_select_add(self._f(x), _dim(d, x.ndim), idx, self._f(v))
    def op_relu(self, x): return self.ops.relu(self._f(x))
    def op_gelu(self, x): return self.ops.gelu(self._f(x))
    def op_sigmoid(self, x): return self.ops.sigmoid(self._f(x))
    def op_softmax(self, x, d): return self.ops.softmax(self._f(x), _dim(d, x.ndim))
    def op_log_softmax(self, x, d): return self.ops.log_softmax(self._f(x), _dim(d, x.ndim))
