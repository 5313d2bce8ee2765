"""GPU parity: float_matmul (path c, tcgen05 GEMM) vs the burn-ndarray oracle.

Protocol of crates/burn-backend-tests/tests/cubecl/matmul.rs:6-78 (random operands, device vs
CPU reference) plus the layout cases of tests/tensor/float/ops/matmul.rs (broadcast batch dims,
transposed views, vec-mat).  Stated tolerances, as a componentwise backward-error bound
|C - C_ref| <= tol * (|A|·|B|) against an f64 reference (and the f32 oracle where it runs):
    F32X3 (3xTF32)  4e-6   — near-f32: what the reference's `Tolerance::rel_abs(1e-5, …)` tests need
    TF32            1e-3   — operands truncated to 10 mantissa bits by the tensor core
    BF16            8e-3   — operands rounded to 8 mantissa bits
Small-integer matrices are exact in every mode (matmul.rs:6-16 asserts equality).
"""
import numpy as np
import pytest

from burn_b200 import _abi as abi
from burn_b200 import ops
from burn_b200.device import DeviceTensor, TapeBuilder
from oracle import oracle
from tests import helpers as H

pytestmark = pytest.mark.gpu

TOL = {abi.MM_F32X3: 4e-6, abi.MM_TF32: 1e-3, abi.MM_BF16: 8e-3}


def rnd(shape, seed):
    return np.random.default_rng(seed).uniform(-0.5, 0.5, shape).astype(np.float32)


def check(got, a, b, precision, what="", tol=None):
    ref = np.matmul(a.astype(np.float64), b.astype(np.float64))
    bound = np.matmul(np.abs(a).astype(np.float64), np.abs(b).astype(np.float64))
    assert got.shape == ref.shape, f"{what}: shape {got.shape} != {ref.shape}"
    err = np.abs(got.astype(np.float64) - ref)
    lim = (TOL[precision] if tol is None else tol) * bound + 1e-30
    if not (err <= lim).all():
        i = np.unravel_index(np.argmax(err / lim), err.shape)
        raise AssertionError(f"{what}: worst at {i}: got {got[i]} ref {ref[i]} err {err[i]:.3e} limit {lim[i]:.3e}")


SHAPES = [(128, 128, 128), (256, 384, 512), (1, 1, 1), (5, 3, 7), (130, 70, 33), (64, 200, 1000), (1000, 40, 96),
          (129, 257, 31),
          # tall/wide enough for the CTA-pair (cta_group::2, 256x256 tile) kernel, with ragged edges
          (1536, 1280, 96), (1300, 1100, 330), (2048, 1024, 64)]


@pytest.mark.parametrize("m,n,k", SHAPES)
@pytest.mark.parametrize("precision", [abi.MM_F32X3, abi.MM_TF32, abi.MM_BF16])
def test_matmul_2d_vs_reference(dev, m, n, k, precision):
    a, b = rnd((m, k), 1), rnd((k, n), 2)
    got = ops.float_matmul(H.up(a), H.up(b), precision).numpy()
    check(got, a, b, precision, f"[{m},{k}]x[{k},{n}]")


def test_f32x3_matches_oracle_at_reference_tolerance(dev):
    # tests/cubecl/matmul.rs uses Tolerance::rel_abs(1e-5 … 1e-4, 1e-5) against ndarray
    a, b = rnd((96, 300), 3), rnd((300, 150), 4)
    got = ops.float_matmul(H.up(a), H.up(b), abi.MM_F32X3).numpy()
    H.assert_close(got, oracle.float_matmul(a[None], b[None])[0], 1e-5, 1e-5, "F32X3 vs oracle")


@pytest.mark.parametrize("precision", [abi.MM_F32X3, abi.MM_TF32, abi.MM_BF16])
def test_small_integers_are_exact(dev, precision):
    rng = np.random.default_rng(5)
    a = rng.integers(-8, 9, (70, 90)).astype(np.float32)
    b = rng.integers(-8, 9, (90, 50)).astype(np.float32)
    got = ops.float_matmul(H.up(a), H.up(b), precision).numpy()
    H.assert_exact(got, (a.astype(np.float64) @ b.astype(np.float64)).astype(np.float32))


@pytest.mark.parametrize("precision", [abi.MM_F32X3, abi.MM_TF32, abi.MM_BF16])
def test_transposed_views_nt_tn_tt(dev, precision):
    # autodiff's matmul backward issues grad·rhsᵀ and lhsᵀ·grad on swap_dims views
    # (crates/burn-autodiff/src/ops/tensor.rs:577-593)
    m, n, k = 136, 200, 72
    a, b = rnd((m, k), 6), rnd((k, n), 7)
    at = H.up(np.ascontiguousarray(a.T)).swap_dims(0, 1)   # [m,k] view with strides (1, m)
    bt = H.up(np.ascontiguousarray(b.T)).swap_dims(0, 1)   # [k,n] view with strides (1, k)
    check(ops.float_matmul(at, H.up(b), precision).numpy(), a, b, precision, "TN")
    check(ops.float_matmul(H.up(a), bt, precision).numpy(), a, b, precision, "NT")
    check(ops.float_matmul(at, bt, precision).numpy(), a, b, precision, "TT")


@pytest.mark.parametrize("precision", [abi.MM_F32X3, abi.MM_TF32])
def test_batched_and_broadcast(dev, precision):
    a = rnd((6, 40, 64), 8)
    b = rnd((6, 64, 48), 9)
    check(ops.float_matmul(H.up(a), H.up(b), precision).numpy(), a, b, precision, "batched")
    b1 = rnd((1, 64, 48), 10)
    check(ops.float_matmul(H.up(a), H.up(b1), precision).numpy(), a, b1, precision, "rhs broadcast")
    a4 = rnd((2, 1, 33, 16), 11)
    b4 = rnd((1, 3, 16, 20), 12)
    check(ops.float_matmul(H.up(a4), H.up(b4), precision).numpy(), a4, b4, precision, "4-D cross broadcast")
    # attention-shaped: [B,H,S,dk] x [B,H,dk,S] with K given as a swap_dims view
    q = rnd((2, 4, 64, 32), 13)
    kk = rnd((2, 4, 64, 32), 14)
    kt = H.up(kk).swap_dims(2, 3)
    check(ops.float_matmul(H.up(q), kt, precision).numpy(), q, kk.swapaxes(2, 3), precision, "q·kᵀ")


def test_bf16_operands_in_place_all_layouts(dev):
    # bf16 storage consumed directly by TMA in its native orientation (K-major and MN-major)
    m, n, k = 192, 136, 160
    a, b = rnd((m, k), 30), rnd((k, n), 31)
    da, db = DeviceTensor.from_bf16_of(a), DeviceTensor.from_bf16_of(b)
    ar, br = da.numpy(), db.numpy()          # the bf16-rounded values, as f32
    dat = DeviceTensor.from_bf16_of(np.ascontiguousarray(a.T)).swap_dims(0, 1)
    dbt = DeviceTensor.from_bf16_of(np.ascontiguousarray(b.T)).swap_dims(0, 1)
    for name, x, y in (("NN", da, db), ("NT", da, dbt), ("TN", dat, db), ("TT", dat, dbt)):
        got = ops.float_matmul(x, y, abi.MM_BF16).numpy()
        ref = ar.astype(np.float64) @ br.astype(np.float64)
        bound = np.abs(ar).astype(np.float64) @ np.abs(br).astype(np.float64)
        assert (np.abs(got - ref) <= 2e-6 * bound + 1e-30).all(), f"bf16 {name}: products are exact, only f32 accumulation error allowed"


def test_linear_shape_with_broadcast_weight(dev):
    # Linear on [B,S,d_in]: Matmul([B,S,d_in]·[1,d_in,d_out])
    x = rnd((4, 32, 96), 15)
    w = rnd((1, 96, 160), 16)
    got = ops.float_matmul(H.up(x), H.up(w), abi.MM_F32X3).numpy()
    check(got, x, w, abi.MM_F32X3, "linear")
    H.assert_close(got, oracle.float_matmul(x, w), 1e-5, 1e-5, "linear vs oracle")


@pytest.mark.parametrize("precision", [abi.MM_F32X3, abi.MM_TF32, abi.MM_BF16])
def test_fused_bias_gelu_epilogue(dev, precision):
    # MatmulOptimization: matmul + fuse-on-write chain gelu(C + bias[N])
    m, n, k = 200, 256, 128
    a, b = rnd((m, k), 17), rnd((k, n), 18)
    bias = rnd((1, n), 19)
    tb = TapeBuilder().op("ADD_F", ("in", 0), ("in", 1), tmp=0)
    H.gelu_tape(tb, ("tmp", 0), out=0)
    got = ops.float_matmul(H.up(a), H.up(b), precision, epilogue=tb.build(),
                           epi_inputs=[H.up(bias)]).numpy()
    plain = ops.float_matmul(H.up(a), H.up(b), precision).numpy()
    want = oracle.gelu(oracle.float_add(plain, bias))
    # same GEMM result in, same op chain: only the erf 1-ulp allowance applies
    H.assert_close(got, want, H.REL_ELEMWISE, H.ABS_GELU, "fused epilogue == unfused chain")


@pytest.mark.parametrize("precision", [abi.MM_F32X3, abi.MM_TF32, abi.MM_BF16])
def test_cta_pair_kernel_layouts_batches_and_epilogue(dev, precision):
    """Shapes that select the CTA-pair kernel: every operand orientation, a broadcast batch, and the
    fused bias+gelu epilogue on a ragged [.., 1100, 1300] output."""
    m, n, k = 1100, 1300, 200
    a, b = rnd((m, k), 31), rnd((k, n), 32)
    da_t = H.up(np.ascontiguousarray(a.T)).swap_dims(0, 1)
    db_t = H.up(np.ascontiguousarray(b.T)).swap_dims(0, 1)
    for da, db, what in ((H.up(a), H.up(b), "NN"), (H.up(a), db_t, "NT"), (da_t, H.up(b), "TN"), (da_t, db_t, "TT")):
        check(ops.float_matmul(da, db, precision).numpy(), a, b, precision, what)
    a3, b3 = rnd((3, 1, 520, 72), 33), rnd((1, 2, 72, 1032), 34)
    check(ops.float_matmul(H.up(a3), H.up(b3), precision).numpy(), a3, b3, precision, "broadcast batch")
    n4 = 1300
    bias = rnd((1, n4), 35)
    tb = TapeBuilder().op("ADD_F", ("in", 0), ("in", 1), tmp=0)
    H.gelu_tape(tb, ("tmp", 0), out=0)
    got = ops.float_matmul(H.up(a), H.up(b), precision, epilogue=tb.build(), epi_inputs=[H.up(bias)]).numpy()
    plain = ops.float_matmul(H.up(a), H.up(b), precision).numpy()
    # erf's 1-ulp allowance scales with |x|/2 in gelu; pre-activations reach |x| ~ 5 here (ABS_GELU covers |x| <= 4)
    H.assert_close(got, oracle.gelu(oracle.float_add(plain, bias)), H.REL_ELEMWISE, 2e-7, "pair epilogue")


@pytest.mark.parametrize("precision", [abi.MM_F32X3, abi.MM_TF32, abi.MM_BF16])
def test_split_k_long_contractions(dev, precision):
    """Few output tiles, long K: the launcher splits K across CTAs / CTA pairs into workspace partials
    and combines them with the deterministic column reduce (weight-gradient shapes, xᵀ·g)."""
    for (m, n, k) in ((1024, 1024, 8192), (256, 384, 4100), (130, 70, 4096)):
        a, b = rnd((m, k), 41), rnd((k, n), 42)
        check(ops.float_matmul(H.up(a), H.up(b), precision).numpy(), a, b, precision, f"split-K NN {m}x{n}x{k}")
    m, n, k = 512, 768, 4096                     # dW = xᵀ·g: both operands MN-major views
    x, g = rnd((k, m), 43), rnd((k, n), 44)
    got = ops.float_matmul(H.up(x).swap_dims(0, 1), H.up(g), precision).numpy()
    check(got, np.ascontiguousarray(x.T), g, precision, "split-K TN")
    a3, b3 = rnd((2, 200, 4096), 45), rnd((2, 4096, 300), 46)
    check(ops.float_matmul(H.up(a3), H.up(b3), precision).numpy(), a3, b3, precision, "split-K batched")
    # determinism: the combine has a fixed order
    a, b = rnd((1024, 8192), 47), rnd((8192, 1024), 48)
    r1 = ops.float_matmul(H.up(a), H.up(b), precision).numpy()
    r2 = ops.float_matmul(H.up(a), H.up(b), precision).numpy()
    assert np.array_equal(r1, r2)


def test_inner_dim_mismatch_is_an_error(dev):
    with pytest.raises(ops.ShapeError):
        ops.float_matmul(H.up(rnd((4, 5), 0)), H.up(rnd((6, 7), 0)))
    with pytest.raises(ops.ShapeError):
        ops.float_matmul(H.up(rnd((2, 4, 5), 0)), H.up(rnd((3, 5, 7), 0)))


def test_large_square_matmul_properties(dev):
    """n = 4096 (BASELINE configs[2] sweep member): sampled rows against f64, plus linearity
    (A·(2B) == 2·(A·B) exactly, scaling by a power of two commutes with every rounding)."""
    n = 4096
    a, b = rnd((n, n), 20), rnd((n, n), 21)
    da, db = H.up(a), H.up(b)
    for precision in (abi.MM_TF32, abi.MM_BF16, abi.MM_F32X3):
        c = ops.float_matmul(da, db, precision).numpy()
        rows = [0, 1, 777, 4095]
        check(c[rows], a[rows], b, precision, f"n=4096 rows, precision {precision}")
        c2 = ops.float_matmul(da, H.up(b * np.float32(2.0)), precision).numpy()
        assert np.array_equal(c2, c * np.float32(2.0))


def _pattern(shape, phase):
    """The reference benches' deterministic generator ((i % 1000) / 1000) - 0.5
    (crates/burn-backend-tests/benches/matmul.rs:30-35), offset by `phase` so lhs != rhs."""
    n = int(np.prod(shape))
    i = (np.arange(n, dtype=np.int64) + phase) % 1000
    return (i.astype(np.float32) / np.float32(1000.0) - np.float32(0.5)).reshape(shape)


@pytest.mark.parametrize("n", [8192, 16384])
def test_configs2_large_squares_all_precisions(dev, n):
    """BASELINE configs[2] members 8192^3 and 16384^3: sampled rows AND sampled columns (so every
    rasterisation group and both CTAs of a pair are hit) against float64, in all three precisions."""
    a = rnd((n, n), 50 + n)
    b = _pattern((n, n), 7)
    da, db = H.up(a), H.up(b)
    rows = [0, 127, 128, 255, 256, 2049, n // 2 + 3, n - 129, n - 1]
    cols = [0, 255, 256, 1023, n // 2 - 1, n - 257, n - 1]
    # F32X3: the products are near-exact, what remains is the tensor core's f32 accumulation over K = n terms
    # (measured 4.2e-6 of |A|·|B| at K = 8192); 2e-5 stays inside the reference's own device-vs-ndarray matmul
    # tolerance (tests/cubecl/matmul.rs:6-78: rel 1e-5 ... 1e-4)
    for precision in (abi.MM_TF32, abi.MM_BF16, abi.MM_F32X3):
        tol = 2e-5 if precision == abi.MM_F32X3 else None
        c = ops.float_matmul(da, db, precision).numpy()
        check(c[rows], a[rows], b, precision, f"n={n} rows, precision {precision}", tol)
        check(c[:, cols], a, b[:, cols], precision, f"n={n} cols, precision {precision}", tol)
        del c


def test_configs2_batched_2048(dev):
    """BASELINE configs[2]: batched [64, 2048, 2048] x [64, 2048, 2048], tf32 and bf16; three whole batch
    members (first, middle, last) against float64."""
    a = rnd((64, 2048, 2048), 60)
    b = rnd((64, 2048, 2048), 61)
    da, db = H.up(a), H.up(b)
    for precision in (abi.MM_TF32, abi.MM_BF16):
        c = ops.float_matmul(da, db, precision).numpy()
        for i in (0, 31, 63):
            check(c[i], a[i], b[i], precision, f"batched member {i}, precision {precision}")
        del c


@pytest.mark.parametrize("precision", [abi.MM_TF32, abi.MM_BF16])
def test_configs2_bias_gelu_epilogue_8192(dev, precision):
    """BASELINE configs[2]: gelu(C + bias[N]) fused into the 8192^3 GEMM == the unfused chain on the plain
    GEMM result (same accumulator in, same op chain: only the erf 1-ulp allowance)."""
    n = 8192
    a, b = rnd((n, n), 70), rnd((n, n), 71)
    bias = rnd((1, n), 72)
    da, db = H.up(a), H.up(b)
    tb = TapeBuilder().op("ADD_F", ("in", 0), ("in", 1), tmp=0)
    H.gelu_tape(tb, ("tmp", 0), out=0)
    got = ops.float_matmul(da, db, precision, epilogue=tb.build(), epi_inputs=[H.up(bias)]).numpy()
    plain = ops.float_matmul(da, db, precision).numpy()
    rows = [0, 255, 256, 4097, n - 1]
    check(plain[rows], a[rows], b, precision, "8192 plain rows")
    # erf's 1-ulp allowance scales with |x|/2 where 1+erf cancels (x in about [-6, 0]): 3 * 2^-24 * ... < 4e-7
    want = oracle.gelu(oracle.float_add(plain[::61], bias))
    H.assert_close(got[::61], want, H.REL_ELEMWISE, 4e-7, "8192 fused bias+gelu epilogue == unfused chain")
