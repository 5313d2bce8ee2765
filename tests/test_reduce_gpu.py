"""GPU parity: reductions (path b) vs the burn-ndarray oracle.

Protocol of crates/burn-backend-tests/tests/cubecl/reduce.rs:10-24 (argmax on every axis
of [2,4,8,16], exact) extended to sum/mean/max/min, fused read/write tapes, ties and NaNs.
Tolerances: indices bit-exact; fp32 sums <= 1e-5 relative (BASELINE.json north_star).
"""
import numpy as np
import pytest

from burn_b200 import _abi as abi
from burn_b200 import device as dv
from burn_b200.device import DeviceTensor, TapeBuilder
from oracle import oracle
from tests import helpers as H

pytestmark = pytest.mark.gpu


def rnd(shape, lo=-1.0, hi=1.0, seed=0):
    return np.random.default_rng(seed).uniform(lo, hi, size=shape).astype(np.float32)


def keep(shape, axis):
    s = list(shape)
    s[axis] = 1
    return tuple(s)


def reduce_axis(kind, x, axis, out_dtype=abi.F32, t=None):
    t = t if t is not None else H.up(x)
    out = DeviceTensor.empty(keep(t.shape, axis), out_dtype)
    dv.launch_reduce(kind, axis, t.shape, [t], [out])
    return out.numpy()


SHAPES = [(2, 4, 8, 16), (37, 100), (64, 256), (300, 1024), (5, 7, 9), (1200, 64), (3, 8192), (8192, 8)]


@pytest.mark.parametrize("shape", SHAPES)
def test_sum_and_mean_every_axis(dev, shape):
    x = rnd(shape)
    t = H.up(x)
    # absolute floor: sums of U(-1,1) cancel, so scale the tolerance by sum|x| like any
    # summation-order bound does (eps * n) — 1e-5 relative to the result otherwise.
    for axis in range(len(shape)):
        s_abs = np.abs(x).sum(axis=axis, keepdims=True).max()
        got = reduce_axis(abi.RED_SUM, x, axis, t=t)
        H.assert_close(got, oracle.float_sum_dim(x, axis), H.REL_REDUCE, 1e-6 * s_abs, f"sum axis {axis}")
        got = reduce_axis(abi.RED_MEAN, x, axis, t=t)
        H.assert_close(got, oracle.float_mean_dim(x, axis), H.REL_REDUCE, 1e-6 * s_abs / shape[axis],
                       f"mean axis {axis}")


@pytest.mark.parametrize("shape", SHAPES)
def test_sum_positive_data_tight_relative(dev, shape):
    x = rnd(shape, 0.0, 1.0, seed=3)
    for axis in range(len(shape)):
        got = reduce_axis(abi.RED_SUM, x, axis)
        H.assert_close(got, oracle.float_sum_dim(x, axis), H.REL_REDUCE, 0.0, f"sum axis {axis}")


@pytest.mark.parametrize("shape", SHAPES)
def test_argmax_argmin_every_axis_exact(dev, shape):
    x = rnd(shape, seed=5)
    t = H.up(x)
    for axis in range(len(shape)):
        for kind, fn in ((abi.RED_ARGMAX, oracle.float_argmax), (abi.RED_ARGMIN, oracle.float_argmin)):
            got = reduce_axis(kind, x, axis, abi.I32, t=t)
            H.assert_exact(got, fn(x, axis), f"arg kind {kind} axis {axis}")
            got64 = reduce_axis(kind, x, axis, abi.I64, t=t)
            H.assert_exact(got64, fn(x, axis), f"arg kind {kind} axis {axis} (i64 out)")


def test_argmax_ties_take_first_and_nan_wins(dev):
    x = np.round(rnd((64, 512), seed=9) * 3)  # many exact ties
    x[3, 100] = np.nan
    x[3, 400] = np.nan  # first NaN (100) must win
    x[7, :] = 1.0
    x[9, 511] = np.nan
    for axis in (0, 1):
        H.assert_exact(reduce_axis(abi.RED_ARGMAX, x, axis, abi.I32), oracle.float_argmax(x, axis), f"argmax {axis}")
        H.assert_exact(reduce_axis(abi.RED_ARGMIN, x, axis, abi.I32), oracle.float_argmin(x, axis), f"argmin {axis}")


def test_arg_on_constant_infinite_rows(dev):
    # all -inf / +inf lanes: the first index must be reported (the oracle folds from (arr[0], 0))
    x = np.full((300, 4096), -np.inf, dtype=np.float32)
    x[5, 100] = 1.0
    for axis in (0, 1):
        H.assert_exact(reduce_axis(abi.RED_ARGMAX, x, axis, abi.I32), oracle.float_argmax(x, axis))
        H.assert_exact(reduce_axis(abi.RED_ARGMIN, -x, axis, abi.I32), oracle.float_argmin(-x, axis))
        H.assert_exact(reduce_axis(abi.RED_MAX, x, axis), oracle.float_max_dim(x, axis))


@pytest.mark.parametrize("shape", [(2, 7, 160000), (1, 16, 1 << 20), (3, 64, 110000)])
def test_short_axis_over_many_columns(dev, shape):
    """[outer, R <= 64, inner] with hundreds of thousands of columns (a split-K combine, per-CTA partial rows): the
    thread-per-column kernel — every kind, ties and NaNs included."""
    x = rnd(shape, seed=21)
    x[0, 3, 5] = np.nan
    x[0, 1, 9] = x[0, 4, 9] = 7.0                                     # a tie: the first index wins
    H.assert_close(reduce_axis(abi.RED_SUM, np.nan_to_num(x), 1), oracle.float_sum_dim(np.nan_to_num(x), 1), H.REL_REDUCE, 1e-6 * shape[1], "sum")
    H.assert_close(reduce_axis(abi.RED_MEAN, np.nan_to_num(x), 1), oracle.float_mean_dim(np.nan_to_num(x), 1), H.REL_REDUCE, 1e-6, "mean")
    H.assert_exact(reduce_axis(abi.RED_MAX, x, 1), oracle.float_max_dim(x, 1), "max")
    H.assert_exact(reduce_axis(abi.RED_MIN, x, 1), oracle.float_min_dim(x, 1), "min")
    H.assert_exact(reduce_axis(abi.RED_ARGMAX, x, 1, abi.I32), oracle.float_argmax(x, 1), "argmax")
    H.assert_exact(reduce_axis(abi.RED_ARGMIN, x, 1, abi.I32), oracle.float_argmin(x, 1), "argmin")


def test_max_min_dim_exact(dev):
    x = rnd((33, 130, 12), seed=2)
    x[1, 5, 3] = np.nan
    for axis in range(3):
        H.assert_exact(reduce_axis(abi.RED_MAX, x, axis), oracle.float_max_dim(x, axis), f"max {axis}")
        H.assert_exact(reduce_axis(abi.RED_MIN, x, axis), oracle.float_min_dim(x, axis), f"min {axis}")


def test_reference_goldens_aggregation(dev):
    # crates/burn-backend-tests/tests/tensor/float/ops/aggregation.rs (sum_dim / mean_dim goldens)
    x = np.array([[0.0, 1.0, 2.0], [3.0, 4.0, 5.0]], dtype=np.float32)
    H.assert_exact(reduce_axis(abi.RED_SUM, x, 1), np.array([[3.0], [12.0]], dtype=np.float32))
    H.assert_exact(reduce_axis(abi.RED_SUM, x, 0), np.array([[3.0, 5.0, 7.0]], dtype=np.float32))
    H.assert_exact(reduce_axis(abi.RED_MEAN, x, 1), np.array([[1.0], [4.0]], dtype=np.float32))
    H.assert_exact(reduce_axis(abi.RED_MEAN, x, 0), np.array([[1.5, 2.5, 3.5]], dtype=np.float32))
    # arg.rs:6-50
    y = np.array([[10.0, 11.0, 2.0], [3.0, 4.0, 5.0]], dtype=np.float32)
    H.assert_exact(reduce_axis(abi.RED_ARGMAX, y, 0, abi.I32), np.array([[0, 0, 1]]))
    H.assert_exact(reduce_axis(abi.RED_ARGMAX, y, 1, abi.I32), np.array([[1], [2]]))
    z = np.array([[10.0, 11.0, 2.0], [30.0, 4.0, 5.0]], dtype=np.float32)
    H.assert_exact(reduce_axis(abi.RED_ARGMIN, z, 0, abi.I32), np.array([[0, 1, 0]]))
    H.assert_exact(reduce_axis(abi.RED_ARGMIN, z, 1, abi.I32), np.array([[2], [1]]))


def test_argmax_on_permuted_view(dev):
    # crates/burn-backend-tests/tests/tensor/float/ops/arg.rs:94-110 (permuted 4-D regression)
    x = np.arange(2 * 3 * 4 * 5, dtype=np.float32).reshape(2, 3, 4, 5)
    t = H.up(x).permute([0, 2, 1, 3])
    xp = np.ascontiguousarray(x.transpose(0, 2, 1, 3))
    for axis in range(4):
        got = reduce_axis(abi.RED_ARGMAX, None, axis, abi.I32, t=t)
        H.assert_exact(got, oracle.float_argmax(xp, axis), f"permuted argmax {axis}")
        got = reduce_axis(abi.RED_SUM, None, axis, t=t)
        H.assert_close(got, oracle.float_sum_dim(xp, axis), H.REL_REDUCE, 0.0, f"permuted sum {axis}")


@pytest.mark.parametrize("n", [1, 5, 1000, 1 << 16, (1 << 20) + 12, 3 * (1 << 20)])
def test_full_sum(dev, n):
    x = rnd((n,), 0.0, 1.0, seed=11)
    out = DeviceTensor.empty((1,))
    dv.launch_reduce_full(abi.RED_SUM, H.up(x), out)
    H.assert_close(out.numpy(), oracle.float_sum(x), H.REL_REDUCE, 0.0, "full sum")
    dv.launch_reduce_full(abi.RED_MEAN, H.up(x), out)
    H.assert_close(out.numpy(), oracle.float_mean(x), H.REL_REDUCE, 0.0, "full mean")


def test_full_sum_of_strided_view(dev):
    x = rnd((96, 200), 0.0, 1.0, seed=4)
    t = H.up(x).swap_dims(0, 1)
    out = DeviceTensor.empty((1,))
    dv.launch_reduce_full(abi.RED_SUM, t, out)
    H.assert_close(out.numpy(), oracle.float_sum(x), H.REL_REDUCE, 0.0)


def test_fused_read_tape_gelu_then_sum(dev):
    # SURVEY §8d-2 fused variant: sum_dim(gelu(a*b), 1)
    a, b = rnd((256, 1024), seed=1), rnd((256, 1024), seed=2)
    tb = TapeBuilder().op("MUL_F", ("in", 0), ("in", 1), tmp=0)
    H.gelu_tape(tb, ("tmp", 0))
    out = DeviceTensor.empty((256, 1))
    dv.launch_reduce(abi.RED_SUM, 1, a.shape, [H.up(a), H.up(b)], [out], read=tb.build())
    want = oracle.float_sum_dim(oracle.gelu(oracle.float_mul(a, b)), 1)
    H.assert_close(out.numpy(), want, H.REL_REDUCE, 1e-6 * 1024, "fused gelu+sum rows")
    out0 = DeviceTensor.empty((1, 1024))
    dv.launch_reduce(abi.RED_SUM, 0, a.shape, [H.up(a), H.up(b)], [out0], read=tb.build())
    want0 = oracle.float_sum_dim(oracle.gelu(oracle.float_mul(a, b)), 0)
    H.assert_close(out0.numpy(), want0, H.REL_REDUCE, 1e-6 * 256, "fused gelu+sum cols")


def test_fused_write_tape(dev):
    # mean_dim followed by (mean + eps).sqrt() * w  — the layer-norm denominator shape of work
    x = rnd((128, 512), 0.0, 2.0, seed=6)
    w = rnd((128, 1), 0.5, 1.5, seed=7)
    wt = (TapeBuilder().op("ADD_F", ("in", 0), ("f", 1e-5))
          .op("SQRT_F", "acc")
          .op("MUL_F", "acc", ("in", 1), out=0))
    out = DeviceTensor.empty((128, 1))
    dv.launch_reduce(abi.RED_MEAN, 1, x.shape, [H.up(x)], [out], write=wt.build(), write_inputs=[H.up(w)])
    want = oracle.float_mul(oracle.float_sqrt(oracle.float_add_scalar(oracle.float_mean_dim(x, 1), 1e-5)), w)
    H.assert_close(out.numpy(), want, H.REL_REDUCE, 0.0)


def test_int_and_bool_reductions(dev):
    rng = np.random.default_rng(0)
    a = rng.integers(-100, 100, size=(40, 65)).astype(np.int32)
    for axis in (0, 1):
        H.assert_exact(reduce_axis(abi.RED_SUM, a, axis, abi.I32), a.sum(axis=axis, keepdims=True))
        H.assert_exact(reduce_axis(abi.RED_MAX, a, axis, abi.I32), a.max(axis=axis, keepdims=True))
        H.assert_exact(reduce_axis(abi.RED_ARGMAX, a, axis, abi.I32), oracle.float_argmax(a, axis))
    m = rng.random((40, 65)) < 0.02
    for axis in (0, 1):
        H.assert_exact(reduce_axis(abi.RED_ANY, m, axis, abi.BOOL), m.any(axis=axis, keepdims=True))
        H.assert_exact(reduce_axis(abi.RED_ALL, ~m, axis, abi.BOOL), (~m).all(axis=axis, keepdims=True))


def test_bad_axis_is_an_error(dev):
    # crates/burn-backend-tests/tests/tensor/float/ops/arg.rs:64-69 (#[should_panic])
    t = H.up(rnd((2, 3)))
    out = DeviceTensor.empty((2, 1), abi.I32)
    with pytest.raises(abi.B200Error):
        dv.launch_reduce(abi.RED_ARGMAX, 2, t.shape, [t], [out])


def test_full_size_reductions(dev):
    """[8192, 8192] (BASELINE configs[1]): sum/mean/argmax on both axes + full sum.  Row sums
    are compared with the oracle (it finishes in ~0.1 s per call); plus the size-independent
    check sum(sum_dim(x, 0)) == sum(sum_dim(x, 1)) == sum(x) within tolerance."""
    n = 8192
    x = rnd((n, n), seed=21)
    t = H.up(x)
    r1 = reduce_axis(abi.RED_SUM, x, 1, t=t)
    r0 = reduce_axis(abi.RED_SUM, x, 0, t=t)
    H.assert_close(r1, oracle.float_sum_dim(x, 1), H.REL_REDUCE, 1e-6 * n)
    H.assert_close(r0, oracle.float_sum_dim(x, 0), H.REL_REDUCE, 1e-6 * n)
    H.assert_exact(reduce_axis(abi.RED_ARGMAX, x, 1, abi.I32, t=t), oracle.float_argmax(x, 1))
    H.assert_exact(reduce_axis(abi.RED_ARGMAX, x, 0, abi.I32, t=t), oracle.float_argmax(x, 0))
    out = DeviceTensor.empty((1,))
    dv.launch_reduce_full(abi.RED_SUM, t, out)
    total = float(out.numpy()[0])
    ref = float(x.astype(np.float64).sum())
    assert abs(total - ref) <= 1e-5 * np.abs(x).sum(dtype=np.float64)
    assert abs(float(r1.astype(np.float64).sum()) - ref) <= 1e-5 * np.abs(x).sum(dtype=np.float64)
    assert abs(float(r0.astype(np.float64).sum()) - ref) <= 1e-5 * np.abs(x).sum(dtype=np.float64)


@pytest.mark.parametrize("shape,axis", [((4096, 1024), 1), ((512, 8192), 1), ((8, 262144), 1), ((2048, 2048), 0),
                                        ((16, 4096, 64), 1), ((3, 1000, 1024), 2)])
@pytest.mark.parametrize("kind", ["sum", "mean", "max", "min"])
def test_large_fuse_on_read_reductions(dev, shape, axis, kind):
    """Sizes above the specialisation threshold (NVRTC kernels from the read tape, jit.cu): rows as a warp,
    as a CTA, split across CTAs, and columns — reduce(gelu(a*b) * 0.5, axis) against the oracle chain."""
    a, b = rnd(shape, seed=11), rnd(shape, seed=12)
    tb = TapeBuilder().op("MUL_F", ("in", 0), ("in", 1), tmp=0)
    H.gelu_tape(tb, ("tmp", 0))
    tb.op("MUL_F", "acc", ("in", 2))
    half = DeviceTensor.from_numpy(np.float32([0.5])).reshape((1,) * len(shape)).expand(shape)   # broadcast operand
    keep = list(shape)
    keep[axis] = 1
    out = DeviceTensor.empty(keep)
    code = {"sum": abi.RED_SUM, "mean": abi.RED_MEAN, "max": abi.RED_MAX, "min": abi.RED_MIN}[kind]
    dv.launch_reduce(code, axis, shape, [H.up(a), H.up(b), half], [out], read=tb.build())
    y = oracle.float_mul_scalar(oracle.gelu(oracle.float_mul(a, b)), 0.5)
    want = {"sum": oracle.float_sum_dim, "mean": oracle.float_mean_dim, "max": oracle.float_max_dim,
            "min": oracle.float_min_dim}[kind](y, axis)
    n = shape[axis]
    abs_tol = {"sum": 2e-7 * n, "mean": 2e-7, "max": 2e-7, "min": 2e-7}[kind]     # erf 1-ulp allowance per element
    H.assert_close(out.numpy(), want, H.REL_REDUCE, abs_tol, f"fused {kind} over axis {axis} of {shape}")


@pytest.mark.parametrize("shape", [(16384, 512), (8192, 1024), (4100, 768), (2, 8192, 256)])
@pytest.mark.parametrize("kind", ["sum", "mean"])
def test_tall_narrow_column_sums(dev, shape, kind):
    """Bias-gradient shapes ([tokens, d_model] summed over the tokens): more row splits than a cluster holds, written as
    partial rows and finished by the short-axis kernel (reduce_fast.cuh, ColParams::partials).  Against the oracle's
    sum_axis / mean_axis (crates/burn-ndarray/src/ops/base.rs sum_dim / mean_dim)."""
    x = rnd(shape, seed=31)
    axis = len(shape) - 2
    code = abi.RED_SUM if kind == "sum" else abi.RED_MEAN
    got = reduce_axis(code, x, axis)
    want = (oracle.float_sum_dim if kind == "sum" else oracle.float_mean_dim)(x, axis)
    n = shape[axis]
    H.assert_close(got, want, H.REL_REDUCE, 1e-6 * n if kind == "sum" else 1e-6, f"{kind} over axis {axis} of {shape}")
