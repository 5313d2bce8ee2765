"""GPU parity: fused elementwise tapes (path a) vs the burn-ndarray oracle.

Mirrors the protocol of crates/burn-backend-tests/tests/cubecl/*.rs — random input,
same op on the device and on the CPU reference, then compare — with the tolerances
BASELINE.json states: bit-exact for int / bool / comparison / select results,
<= 1e-6 relative for fp32 elementwise.
"""
import numpy as np
import pytest

from burn_b200 import _abi as abi
from burn_b200.device import DeviceTensor, TapeBuilder
from oracle import oracle
from tests import helpers as H

pytestmark = pytest.mark.gpu

RNG = np.random.default_rng(0)


def rnd(shape, lo=-1.0, hi=1.0, seed=None):
    r = np.random.default_rng(seed) if seed is not None else RNG
    return r.uniform(lo, hi, size=shape).astype(np.float32)


SHAPES = [(7,), (1024,), (33, 65), (4, 8, 16), (2, 3, 4, 5), (256, 1024)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("op,fn", [("ADD_F", oracle.float_add), ("SUB_F", oracle.float_sub),
                                   ("MUL_F", oracle.float_mul), ("DIV_F", oracle.float_div)])
def test_binary_same_shape_is_bit_exact(dev, shape, op, fn):
    a, b = rnd(shape), rnd(shape, 0.5, 2.0)
    got = H.binary(op, H.up(a), H.up(b))
    H.assert_exact(got, fn(a, b), op)  # IEEE add/sub/mul/div: identical bits


@pytest.mark.parametrize("sa,sb", [((4, 1), (1, 8)), ((2, 3, 4), (1, 3, 1)), ((64, 256), (1, 256)),
                                   ((64, 256), (64, 1)), ((1,), (5,)), ((2, 1, 1, 8), (2, 4, 4, 8))])
def test_binary_broadcast(dev, sa, sb):
    a, b = rnd(sa), rnd(sb)
    H.assert_exact(H.binary("ADD_F", H.up(a), H.up(b)), oracle.float_add(a, b), "add broadcast")
    H.assert_exact(H.binary("MUL_F", H.up(a), H.up(b)), oracle.float_mul(a, b), "mul broadcast")


def test_tensor_remainder_is_the_f64_floor_form(dev):
    """Tensor-tensor `remainder` is a - b*floor(a/b) evaluated in f64 (crates/burn-ndarray/src/ops/base.rs:909-922);
    `remainder_scalar` is ((x % y) + y) % y (base.rs:924-930).  The two differ for infinite divisors, in the sign of
    zero and in rounding — separate opcodes (REMT_F / REM_F), each bit-exact against the oracle; the literals of
    crates/burn-backend-tests/tests/tensor/float/ops/remainder.rs:7-19,35-52 on top."""
    a = np.concatenate([rnd((2000,), -20, 20), np.float32([5.0, -5.0, 7.5, -7.5, 0.0, -0.0, 1e30, 3.0, -3.0, 6.0])])
    b = np.concatenate([rnd((2000,), 0.25, 4.0) * np.where(np.arange(2000) % 3 == 0, -1, 1).astype(np.float32),
                        np.float32([np.inf, np.inf, -np.inf, 2.5, 3.0, 3.0, 7.0, -3.0, 3.0, -1.5])])
    with np.errstate(all="ignore"):
        want = oracle.float_remainder(a, b)
        got = H.binary("REMT_F", H.up(a), H.up(b))
        H.assert_exact(got, want, "tensor remainder")
        H.assert_exact(H.binary_scalar("REM_F", H.up(a), -1.5), oracle.float_remainder_scalar(a, -1.5), "scalar remainder")
    lhs, rhs = np.float32([-3.0, -2.0, -1.0, 1.0, 2.0, 3.0]), np.float32([2.0, 3.0, 1.0, 2.0, 1.0, 3.0])
    H.assert_exact(H.binary("REMT_F", H.up(lhs), H.up(rhs)), np.float32([1.0, 1.0, 0.0, 1.0, 0.0, 0.0]))   # remainder.rs:7-19


def test_scalar_ops(dev):
    a = rnd((5, 40))
    H.assert_exact(H.binary_scalar("ADD_F", H.up(a), 2.5), oracle.float_add_scalar(a, 2.5))
    H.assert_exact(H.binary_scalar("MUL_F", H.up(a), -3.0), oracle.float_mul_scalar(a, -3.0))
    H.assert_exact(H.binary_scalar("DIV_F", H.up(a), 1.4142135623730951),
                   oracle.float_div_scalar(a, 1.4142135623730951))
    H.assert_exact(H.binary_scalar("SUB_F", H.up(a), 0.1), oracle.float_sub_scalar(a, 0.1))


@pytest.mark.parametrize("op,fn,lo,hi", [
    ("EXP_F", oracle.float_exp, -10, 10), ("LOG_F", oracle.float_log, 1e-3, 100),
    ("LOG1P_F", oracle.float_log1p, -0.9, 50), ("SQRT_F", oracle.float_sqrt, 0, 100),
    ("TANH_F", oracle.float_tanh, -6, 6), ("ERF_F", oracle.float_erf, -4, 4),
    ("SIN_F", oracle.float_sin, -20, 20), ("COS_F", oracle.float_cos, -20, 20),
    ("RECIP_F", oracle.float_recip, 0.1, 10), ("ABS_F", oracle.float_abs, -5, 5),
    ("NEG_F", oracle.float_neg, -5, 5), ("FLOOR_F", oracle.float_floor, -5, 5),
    ("CEIL_F", oracle.float_ceil, -5, 5), ("ROUND_F", oracle.float_round, -5, 5),
    ("TRUNC_F", oracle.float_trunc, -5, 5), ("SIGN_F", oracle.float_sign, -5, 5),
])
def test_unary(dev, op, fn, lo, hi):
    a = rnd((37, 129), lo, hi)
    got = H.unary(op, H.up(a))
    H.assert_close(got, fn(a), H.REL_ELEMWISE, 0.0, op)


def test_erf_and_tanh_within_one_ulp_of_oracle(dev):
    # The oracle evaluates erf / tanh in f64 and rounds.  Both are f32 routines here, designed to land within 1 ulp of
    # that correctly rounded value (scripts/fit_erf.py, scripts/fit_tanh.py): <= 1 ulp everywhere and identical for
    # > 97 % of inputs (f64 on the device cost the gelu chains a large part of the HBM roofline).
    a = np.concatenate([rnd((1 << 18,), -5, 5), rnd((1 << 18,), -0.75, 0.75), rnd((1 << 16,), -12, 12),
                        np.array([0.0, -0.0, 0.5, -0.5, 0.75, -0.75, 0.7499999, 4.0, -4.0, 3.9999998, 9.125, 9.1249, -9.2, 1e-30,
                                  -1e-38, np.inf, -np.inf], dtype=np.float32)])
    got = H.unary("ERF_F", H.up(a))
    H.assert_ulp(got, oracle.float_erf(a), 1, 0.03, "erf")
    assert np.isnan(H.unary("ERF_F", H.up(np.array([np.nan], dtype=np.float32)))[0])
    H.assert_ulp(H.unary("TANH_F", H.up(a)), oracle.float_tanh(a), 1, 0.03, "tanh")
    assert np.isnan(H.unary("TANH_F", H.up(np.array([np.nan], dtype=np.float32)))[0])
    # mixed groups: lanes inside and outside the polynomial region in the same warp, and a launch that is entirely inside
    mix = np.tile(np.array([0.1, 2.0, -0.3, -3.5], dtype=np.float32), 4096)
    H.assert_ulp(H.unary("ERF_F", H.up(mix)), oracle.float_erf(mix), 1, 0.03, "erf (mixed warps)")
    inner = rnd((1 << 16,), -0.74, 0.74)
    H.assert_ulp(H.unary("ERF_F", H.up(inner)), oracle.float_erf(inner), 1, 0.05, "erf (polynomial region only)")


def test_special_values(dev):
    a = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, 3.4e38, -3.4e38], dtype=np.float32)
    b = np.array([1.0, 0.0, 2.0, np.inf, 1.0, 1e-45, 3.4e38, 1.0], dtype=np.float32)
    with np.errstate(all="ignore"):
        H.assert_exact(H.binary("ADD_F", H.up(a), H.up(b)), oracle.float_add(a, b))
        H.assert_exact(H.binary("MUL_F", H.up(a), H.up(b)), oracle.float_mul(a, b))
        H.assert_exact(H.binary("DIV_F", H.up(a), H.up(b)), oracle.float_div(a, b))


@pytest.mark.parametrize("op,fn", [("EQ_F", oracle.float_equal), ("NE_F", oracle.float_not_equal),
                                   ("LT_F", oracle.float_lower), ("LE_F", oracle.float_lower_equal),
                                   ("GT_F", oracle.float_greater), ("GE_F", oracle.float_greater_equal)])
def test_comparisons_bit_exact(dev, op, fn):
    a = np.round(rnd((31, 64)) * 4) / 4
    b = np.round(rnd((31, 64)) * 4) / 4
    a[0, :4] = [np.nan, 1.0, np.inf, -0.0]
    b[0, :4] = [np.nan, np.nan, np.inf, 0.0]
    got = H.binary(op, H.up(a), H.up(b), out_dtype=abi.BOOL)
    H.assert_exact(got, fn(a, b), op)
    got_s = H.binary_scalar(op, H.up(a), 0.25, out_dtype=abi.BOOL)
    H.assert_exact(got_s, fn(a, np.float32(0.25)), op + " scalar")


def test_mask_fill_and_where(dev):
    a, src = rnd((6, 10, 32)), rnd((6, 10, 32))
    m = RNG.random((6, 10, 32)) < 0.3
    tb = TapeBuilder().op("SELECT", ("in", 0), ("f", 2.0), ("in", 1), out=0)
    H.assert_exact(H.run_tape(tb, [H.up(a), H.up(m)], a.shape), oracle.float_mask_fill(a, m, 2.0))
    tb = TapeBuilder().op("SELECT", ("in", 0), ("in", 1), ("in", 2), out=0)
    H.assert_exact(H.run_tape(tb, [H.up(a), H.up(src), H.up(m)], a.shape), oracle.float_mask_where(a, m, src))
    # broadcast pad-mask [B,1,1,S] as MHA builds it (crates/burn-nn/src/modules/attention/mha.rs:282-285)
    x = rnd((2, 4, 8, 16))
    pm = RNG.random((2, 1, 1, 16)) < 0.5
    tb = TapeBuilder().op("SELECT", ("in", 0), ("f", -1.0e9), ("in", 1), out=0)
    H.assert_exact(H.run_tape(tb, [H.up(x), H.up(pm).expand(x.shape)], x.shape),
                   oracle.float_mask_fill(x, pm, -1.0e9))


def test_relu_chain(dev):
    x = rnd((100, 33))
    tb = (TapeBuilder().op("LE_F", ("in", 0), ("f", 0.0), tmp=0)
          .op("SELECT", ("in", 0), ("f", 0.0), ("tmp", 0), out=0))
    H.assert_exact(H.run_tape(tb, [H.up(x)], x.shape), oracle.relu(x))


@pytest.mark.parametrize("shape", [(10,), (64, 128), (3, 5, 7)])
def test_gelu_fused_chain(dev, shape):
    x = rnd(shape, -4, 4)
    tb = H.gelu_tape(TapeBuilder(), ("in", 0), out=0)
    got = H.run_tape(tb, [H.up(x)], shape)
    # 1 + erf cancels for negative x: one f32 ulp of erf (6e-8) is the absolute floor
    H.assert_close(got, oracle.gelu(x), H.REL_ELEMWISE, H.ABS_GELU, "gelu")


def bench_chain_tape():
    """y = mask_fill(gelu(a*b + c), m, 0)  — BASELINE.json configs[1] / SURVEY §8d-1."""
    tb = TapeBuilder()
    tb.op("MUL_F", ("in", 0), ("in", 1))
    tb.op("ADD_F", "acc", ("in", 2), tmp=0)
    H.gelu_tape(tb, ("tmp", 0))
    tb.op("SELECT", "acc", ("f", 0.0), ("in", 3), out=0)
    return tb


@pytest.mark.parametrize("shape", [(128, 256), (1000, 1000), (1024, 1024)])
def test_bench_chain_matches_oracle(dev, shape):
    a, b, c = rnd(shape, seed=0), rnd(shape, seed=1), rnd(shape, seed=2)
    m = a < 0
    want = oracle.float_mask_fill(oracle.gelu(oracle.float_add(oracle.float_mul(a, b), c)), m, 0.0)
    got = H.run_tape(bench_chain_tape(), [H.up(a), H.up(b), H.up(c), H.up(m)], shape)
    H.assert_close(got, want, H.REL_ELEMWISE, 0.0, "bench chain")
    np.testing.assert_array_equal(want, oracle.bench_chain_unfused(a, b, c, m))


def test_bench_chain_variants_broadcast_and_transposed(dev):
    shape = (256, 512)
    a, b = rnd(shape, seed=0), rnd(shape, seed=1)
    c_row = rnd((1, shape[1]), seed=2)
    m = a < 0
    want = oracle.float_mask_fill(oracle.gelu(oracle.float_add(oracle.float_mul(a, b), c_row)), m, 0.0)
    got = H.run_tape(bench_chain_tape(), [H.up(a), H.up(b), H.up(c_row).expand(shape), H.up(m)], shape)
    H.assert_close(got, want, H.REL_ELEMWISE, 0.0, "chain with broadcast c")
    # `a` given as a transposed view of a [512, 256] buffer
    at = H.up(np.ascontiguousarray(a.T)).swap_dims(0, 1)
    got = H.run_tape(bench_chain_tape(), [at, H.up(b), H.up(c_row).expand(shape), H.up(m)], shape)
    H.assert_close(got, want, H.REL_ELEMWISE, 0.0, "chain with transposed a")


def test_multiple_outputs_and_temps(dev):
    x, y = rnd((50, 20)), rnd((50, 20))
    tb = (TapeBuilder().op("ADD_F", ("in", 0), ("in", 1), tmp=0, out=0)
          .op("MUL_F", "acc", ("in", 0), tmp=1)
          .op("SUB_F", ("tmp", 0), ("tmp", 1), out=1)
          .op("EXP_F", ("tmp", 0), out=2))
    o0, o1, o2 = H.run_tape(tb, [H.up(x), H.up(y)], x.shape, (abi.F32, abi.F32, abi.F32))
    s = oracle.float_add(x, y)
    H.assert_exact(o0, s)
    H.assert_exact(o1, oracle.float_sub(s, oracle.float_mul(s, x)))
    H.assert_close(o2, oracle.float_exp(s), H.REL_ELEMWISE)


def test_inplace_output_aliases_input(dev):
    x = rnd((300, 7))
    t = H.up(x)
    from burn_b200 import device as dv
    dv.launch_elemwise(TapeBuilder().op("MUL_F", ("in", 0), ("f", 2.0), out=0).build(), [t], [t], x.shape)
    H.assert_exact(t.numpy(), oracle.float_mul_scalar(x, 2.0))


def test_int_and_bool_ops_exact(dev):
    a = RNG.integers(-1000, 1000, size=(17, 24)).astype(np.int32)
    b = RNG.integers(1, 50, size=(17, 24)).astype(np.int32)
    i32 = abi.I32
    H.assert_exact(H.binary("ADD_I", H.up(a), H.up(b), i32), a + b)
    H.assert_exact(H.binary("MUL_I", H.up(a), H.up(b), i32), a * b)
    H.assert_exact(H.binary("DIV_I", H.up(a), H.up(b), i32), (np.trunc(a / b)).astype(np.int32))
    H.assert_exact(H.binary("AND_I", H.up(a), H.up(b), i32), a & b)
    H.assert_exact(H.binary("XOR_I", H.up(a), H.up(b), i32), a ^ b)
    H.assert_exact(H.binary("LT_I", H.up(a), H.up(b), abi.BOOL), a < b)
    p, q = RNG.random((9, 9)) < 0.5, RNG.random((9, 9)) < 0.5
    H.assert_exact(H.binary("AND_B", H.up(p), H.up(q), abi.BOOL), p & q)
    H.assert_exact(H.binary("OR_B", H.up(p), H.up(q), abi.BOOL), p | q)
    H.assert_exact(H.unary("NOT_B", H.up(p), abi.BOOL), ~p)
    # i64 storage
    a64 = a.astype(np.int64)
    H.assert_exact(H.binary("SUB_I", H.up(a64), H.up(b.astype(np.int64)), abi.I64), a64 - b)


def test_casts(dev):
    x = rnd((8, 33), -100, 100)
    H.assert_exact(H.unary("F2I", H.up(x), abi.I32), x.astype(np.int32))  # trunc toward zero
    i = RNG.integers(-5, 5, size=(8, 33)).astype(np.int32)
    H.assert_exact(H.unary("I2F", H.up(i), abi.F32), i.astype(np.float32))
    m = RNG.random((8, 33)) < 0.5
    H.assert_exact(H.unary("B2F", H.up(m), abi.F32), m.astype(np.float32))
    H.assert_exact(H.unary("B2I", H.up(m), abi.I64), m.astype(np.int64))


def test_half_precision_storage(dev):
    x, y = rnd((64, 96)), rnd((64, 96))
    xh, yh = x.astype(np.float16), y.astype(np.float16)
    got = H.binary("ADD_F", H.up(xh), H.up(yh), abi.F16)
    want = (xh.astype(np.float32) + yh.astype(np.float32)).astype(np.float16)
    H.assert_exact(got, want, "f16 add (f32 math, RN store)")
    xb = DeviceTensor.from_bf16_of(x)
    yb = DeviceTensor.from_bf16_of(y)
    got = H.binary("MUL_F", xb, yb, abi.F32)
    H.assert_exact(got, xb.numpy() * yb.numpy(), "bf16 in, f32 out")


def test_empty_tensor_is_a_noop(dev):
    t = DeviceTensor.empty((0, 8))
    from burn_b200 import device as dv
    dv.launch_elemwise(TapeBuilder().op("NEG_F", ("in", 0), out=0).build(), [t], [t], (0, 8))


def test_shape_mismatch_is_an_error(dev):
    a, b = H.up(rnd((4, 5))), H.up(rnd((4, 6)))
    out = DeviceTensor.empty((4, 5))
    from burn_b200 import device as dv
    with pytest.raises(abi.B200Error) as e:
        dv.launch_elemwise(TapeBuilder().op("ADD_F", ("in", 0), ("in", 1), out=0).build(), [a, b], [out], (4, 5))
    assert e.value.status == abi.ERR_SHAPE


def test_full_size_chain_properties(dev):
    """At the BASELINE size [8192, 8192] the oracle is too slow to run everywhere, so check
    size-independent properties: masked positions are exactly 0, a sampled block equals
    the oracle, and the result is idempotent across launches."""
    n = 8192
    rng = np.random.default_rng(7)
    a = rng.uniform(-1, 1, (n, n)).astype(np.float32)
    b = rng.uniform(-1, 1, (n, n)).astype(np.float32)
    c = rng.uniform(-1, 1, (n, n)).astype(np.float32)
    m = a < 0
    da, db, dc, dm = H.up(a), H.up(b), H.up(c), H.up(m)
    got = H.run_tape(bench_chain_tape(), [da, db, dc, dm], (n, n))
    assert np.all(got[m] == 0.0)
    blk = (slice(4000, 4256), slice(0, 8192))
    want = oracle.float_mask_fill(oracle.gelu(oracle.float_add(oracle.float_mul(a[blk], b[blk]), c[blk])), m[blk], 0.0)
    H.assert_close(got[blk], want, H.REL_ELEMWISE, 0.0, "sampled block")
    again = H.run_tape(bench_chain_tape(), [da, db, dc, dm], (n, n))
    assert np.array_equal(got, again)


def test_division_by_scalar_keeps_the_sign_of_zero(dev):
    """x / c runs as Markstein's exact sequence; -0.0 / 3 must stay -0.0 like IEEE division (oracle: ndarray `/`)."""
    x = np.zeros(64, dtype=np.float32)
    x[::2] = -0.0
    x[5] = 7.5
    from burn_b200 import ops
    for c in (3.0, -3.0, 1.4142135623730951):
        got = ops.float_div_scalar(H.up(x), c).numpy()
        want = x / np.float32(c)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), c
