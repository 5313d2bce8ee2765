"""Runs the reference's golden vectors (tests/golden/burn_backend_tests.json) against a backend:
the CPU oracle (pins the oracle, `-m "not gpu"`) or the CUDA library through the C ABI (`-m gpu`).
"""
from __future__ import annotations

import ctypes as C
import json
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden" / "burn_backend_tests.json"
UNARY = ("erf", "exp", "log", "sqrt", "abs", "ceil", "floor", "neg", "recip", "round", "sign", "log1p", "tanh")


def load_cases():
    return json.loads(GOLDEN.read_text())["cases"]


def to_array(x, dtype=np.float32):
    def conv(v):
        if isinstance(v, list):
            return [conv(e) for e in v]
        if isinstance(v, str):
            return {"nan": np.nan, "inf": np.inf, "-inf": -np.inf}[v]
        return v
    return np.array(conv(x), dtype=dtype)


def apply_view_np(a, view):
    if view is None:
        return a
    if view == "transpose":
        return np.swapaxes(a, -1, -2)
    if view == "swap02":
        return np.swapaxes(a, 0, 2)
    raise ValueError(view)


def check(case, got):
    want = to_array(case["expected"], np.float64)
    got = np.asarray(got)
    assert got.shape == want.shape, f"{case['name']} ({case['cite']}): shape {got.shape} != {want.shape}"
    if case["tol"] == "exact":
        # TensorData::assert_eq(strict=false) converts the expected literals to the tensor's dtype
        w = want.astype(got.dtype) if got.dtype.kind == "f" else want
        same = (got == w) | (np.isnan(got.astype(np.float64)) & np.isnan(want))
        assert same.all(), f"{case['name']} ({case['cite']}): got {got.tolist()} want {want.tolist()}"
    else:
        from oracle import oracle
        rel, abs_ = case["tol"]
        ok = oracle.approx_eq_mask(got, want, rel, abs_)
        assert ok.all(), f"{case['name']} ({case['cite']}): got {got.tolist()} want {want.tolist()}"


# ------------------------------------------------------------------ oracle backend
def run_oracle(case):
    from oracle import oracle as o
    op, args = case["op"], case.get("args", {})
    views = case.get("views", [None] * len(case["inputs"]))
    ins = case["inputs"]

    def f(i):
        return np.ascontiguousarray(apply_view_np(to_array(ins[i]), views[i] if i < len(views) else None))

    if op in ("add", "sub", "mul", "div"):
        return getattr(o, f"float_{op}")(f(0), f(1))
    if op.endswith("_scalar"):
        return getattr(o, f"float_{op}")(f(0), args["scalar"])
    if op in UNARY:
        return getattr(o, f"float_{op}")(f(0))
    if op in ("gelu", "relu", "sigmoid"):
        return getattr(o, op)(f(0))
    if op == "softmax":
        return o.softmax(f(0), args["dim"])
    if op == "log_softmax":
        return o.log_softmax(f(0), args["dim"])
    if op == "clamp":
        return o.float_clamp(f(0), args["min"], args["max"])
    if op == "mask_where":
        return o.float_mask_where(f(0), to_array(ins[1], bool), f(2))
    if op == "mask_fill_le":
        x = f(0)
        return o.float_mask_fill(x, o.float_lower_equal(x, args["le"]), args["value"])
    if op == "mean":
        return o.float_mean(f(0))
    if op == "sum":
        return o.float_sum(f(0))
    if op in ("sum_dim", "mean_dim", "argmax", "argmin"):
        return getattr(o, f"float_{op}")(f(0), args["dim"])
    if op == "matmul":
        return o.float_matmul(f(0), f(1))
    if op == "gather":
        return o.float_gather(args["dim"], f(0), to_array(ins[1], np.int64))
    if op == "scatter_add":
        return o.float_scatter_add(args["dim"], f(0), to_array(ins[1], np.int64), f(2))
    if op == "select":
        return o.float_select(f(0), args["dim"], to_array(ins[1], np.int64))
    if op == "select_add":
        return o.float_select_add(f(0), args["dim"], to_array(ins[1], np.int64), f(2))
    raise NotImplementedError(op)


# ------------------------------------------------------------------ device backend (C ABI)
def run_device(case):
    from burn_b200 import _abi as abi
    from burn_b200 import device as dv
    from burn_b200 import ops
    from burn_b200.device import DeviceTensor

    op, args = case["op"], case.get("args", {})
    views = case.get("views", [None] * len(case["inputs"]))
    ins = case["inputs"]

    def t(i, dtype=np.float32):
        d = DeviceTensor.from_numpy(to_array(ins[i], dtype))
        v = views[i] if i < len(views) else None
        if v == "transpose":
            d = d.swap_dims(d.ndim - 1, d.ndim - 2)
        elif v == "swap02":
            d = d.swap_dims(0, 2)
        return d

    if op in ("add", "sub", "mul", "div"):
        return getattr(ops, f"float_{op}")(t(0), t(1)).numpy()
    if op.endswith("_scalar"):
        return getattr(ops, f"float_{op}")(t(0), args["scalar"]).numpy()
    if op in UNARY:
        return getattr(ops, f"float_{op}")(t(0)).numpy()
    if op in ("gelu", "relu", "sigmoid"):
        return getattr(ops, op)(t(0)).numpy()
    if op == "softmax":
        return ops.softmax(t(0), args["dim"]).numpy()
    if op == "log_softmax":
        return ops.log_softmax(t(0), args["dim"]).numpy()
    if op == "clamp":
        return ops.float_clamp(t(0), args["min"], args["max"]).numpy()
    if op == "mask_where":
        return ops.float_mask_where(t(0), t(1, bool), t(2)).numpy()
    if op == "mask_fill_le":
        x = t(0)
        return ops.float_mask_fill(x, ops.float_lower_equal_elem(x, args["le"]), args["value"]).numpy()
    if op == "mean":
        return ops.float_mean(t(0)).numpy()
    if op == "sum":
        return ops.float_sum(t(0)).numpy()
    if op in ("sum_dim", "mean_dim"):
        return getattr(ops, f"float_{op}")(t(0), args["dim"]).numpy()
    if op in ("argmax", "argmin"):
        return getattr(ops, f"float_{op}")(t(0), args["dim"]).numpy()
    if op == "matmul":
        return ops.float_matmul(t(0), t(1)).numpy()
    if op == "gather":
        return ops.float_gather(args["dim"], t(0), t(1, np.int64)).numpy()
    if op == "scatter_add":
        return ops.float_scatter_add(args["dim"], t(0), t(1, np.int64), t(2)).numpy()
    if op == "select":
        return ops.float_select(t(0), args["dim"], t(1, np.int64)).numpy()
    if op == "select_add":
        return ops.float_select_add(t(0), args["dim"], t(1, np.int64), t(2)).numpy()
    raise NotImplementedError(op)
