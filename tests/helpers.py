"""Shared helpers for the GPU parity tests: run single ops / fused chains through
the C ABI and compare with the CPU oracle."""
from __future__ import annotations

import numpy as np

from burn_b200 import _abi as abi
from burn_b200 import device as dv
from burn_b200.device import DeviceTensor, TapeBuilder

# parity tolerances stated by BASELINE.json north_star.  The comparison is burn's Tolerance helper
# |x-y| < max(R*|x+y|, A) (crates/burn-std/src/data/compare.rs:10-27): relative to |x+y| ~ 2|x|, so the
# north_star's "<= 1e-6 relative error" is R = 5e-7 in that helper's terms.
REL_ELEMWISE = 5e-7
REL_REDUCE = 5e-6
# gelu = x*(1+erf(x/sqrt2))/2: for negative x the 1+erf cancels, so one f32 ulp of erf
# (2^-24 = 6e-8, times |x|/2 <= 2) is the absolute floor of any f32 implementation
ABS_GELU = 1.2e-7


def up(a, dtype=None) -> DeviceTensor:
    return DeviceTensor.from_numpy(np.asarray(a), dtype=dtype)


def run_tape(tb: TapeBuilder, inputs, out_shape, out_dtypes=(abi.F32,)):
    outs = [DeviceTensor.empty(out_shape, dt) for dt in out_dtypes]
    dv.launch_elemwise(tb.build(), inputs, outs, out_shape)
    res = [o.numpy() for o in outs]
    return res[0] if len(res) == 1 else res


def binary(opname: str, a: DeviceTensor, b: DeviceTensor, out_dtype=abi.F32):
    shape = np.broadcast_shapes(a.shape, b.shape)
    tb = TapeBuilder().op(opname, ("in", 0), ("in", 1), out=0)
    return run_tape(tb, [a.expand(shape), b.expand(shape)], shape, (out_dtype,))


def binary_scalar(opname: str, a: DeviceTensor, s, kind="f", out_dtype=abi.F32):
    tb = TapeBuilder().op(opname, ("in", 0), (kind, s), out=0)
    return run_tape(tb, [a], a.shape, (out_dtype,))


def unary(opname: str, a: DeviceTensor, out_dtype=abi.F32):
    tb = TapeBuilder().op(opname, ("in", 0), out=0)
    return run_tape(tb, [a], a.shape, (out_dtype,))


def gelu_tape(tb: TapeBuilder, src, out=None):
    """gelu(x) = x * (erf(x / sqrt2) + 1) / 2 as the 5 primitive ops burn-fusion sees
    (crates/burn-backend/src/backend/ops/activation.rs:69-76).  `src` must be re-readable
    (an input or a temp)."""
    tb.op("DIV_F", src, ("f", 1.4142135623730951))
    tb.op("ERF_F", "acc")
    tb.op("ADD_F", "acc", ("f", 1.0))
    tb.op("MUL_F", src, "acc")
    tb.op("DIV_F", "acc", ("f", 2.0), out=out)
    return tb


def assert_close(got, want, rel, abs_=0.0, what=""):
    from oracle import oracle
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    ok = oracle.approx_eq_mask(got, want, rel, abs_)
    if not ok.all():
        bad = np.argwhere(~ok)
        i = tuple(bad[0])
        raise AssertionError(
            f"{what}: {bad.shape[0]} / {ok.size} elements differ beyond rel={rel} abs={abs_}; "
            f"first at {i}: got {got[i]!r} want {want[i]!r}")


def assert_exact(got, want, what=""):
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    if got.dtype.kind == "f":
        same = (got == want) | (np.isnan(got) & np.isnan(want))
    else:
        same = got.astype(np.int64) == want.astype(np.int64)
    if not same.all():
        i = tuple(np.argwhere(~same)[0])
        raise AssertionError(f"{what}: {np.count_nonzero(~same)} mismatches; first at {i}: got {got[i]!r} want {want[i]!r}")


def assert_ulp(got, want, max_ulp: int, max_mismatch_frac: float, what=""):
    """f32 results within `max_ulp` units in the last place, and bit-identical for all but
    `max_mismatch_frac` of the elements."""
    got = np.asarray(got, dtype=np.float32)
    want = np.asarray(want, dtype=np.float32)
    assert got.shape == want.shape

    def ordered(a):
        i = a.view(np.int32).astype(np.int64)
        return np.where(i < 0, -(i & 0x7FFFFFFF), i)

    fin = np.isfinite(want)
    assert np.array_equal(got[~fin], want[~fin], equal_nan=True), f"{what}: non-finite values differ"
    d = np.abs(ordered(got[fin]) - ordered(want[fin]))
    assert d.max(initial=0) <= max_ulp, f"{what}: max ulp distance {d.max()} > {max_ulp}"
    frac = np.count_nonzero(d) / max(d.size, 1)
    assert frac <= max_mismatch_frac, f"{what}: {frac:.4%} of elements differ (limit {max_mismatch_frac:.2%})"
