"""GPU parity: row-resident softmax / log_softmax / layer_norm (the ReduceBroadcasted analogue,
crates/burn-cubecl-fusion/src/optim/reduce_broadcasted/) vs the oracle's op-by-op chains
(activation.rs:250-276, ops/modules/base.rs:846-877) and vs our own unfused op chain.
Tolerance: 1e-5 relative (row sums are reassociated), tiny absolute floor for softmax tails."""
import numpy as np
import pytest

from burn_b200 import ops
from burn_b200.device import DeviceTensor
from oracle import oracle
from tests import helpers as H

pytestmark = pytest.mark.gpu


def rnd(shape, seed, lo=-3.0, hi=3.0):
    return np.random.default_rng(seed).uniform(lo, hi, shape).astype(np.float32)


SHAPES = [(7, 4), (64, 128), (33, 256), (16, 8, 512), (4, 2, 3, 1024), (9, 2048), (5, 4096), (3, 50257), (2, 1000)]


@pytest.mark.parametrize("shape", SHAPES)
def test_softmax_and_log_softmax_rows(dev, shape):
    x = rnd(shape, 1)
    last = len(shape) - 1
    H.assert_close(ops.softmax_rows(H.up(x)).numpy(), oracle.softmax(x, last), 1e-5, 1e-10, "softmax")
    H.assert_close(ops.softmax_rows(H.up(x), log=True).numpy(), oracle.log_softmax(x, last), 1e-5, 1e-6, "log_softmax")


def test_softmax_reference_goldens(dev):
    # crates/burn-backend-tests/tests/tensor/float/activation/softmax.rs:6,42
    x = np.array([[1.0, 7.0], [13.0, -3.0]], dtype=np.float32)
    H.assert_close(ops.softmax_rows(H.up(x)).numpy(),
                   np.array([[2.472623e-03, 9.975274e-01], [1.0, 1.125352e-07]], dtype=np.float32), 5e-3, 1e-5)
    x = np.array([[-1.0, 0.0, 1.0, 2.0], [0.5, 0.5, 0.5, 0.5]], dtype=np.float32)
    H.assert_close(ops.softmax_rows(H.up(x)).numpy(),
                   np.array([[0.03205860, 0.08714432, 0.23688284, 0.64391422], [0.25] * 4], dtype=np.float32), 5e-3, 1e-5)


def test_fused_softmax_equals_unfused_chain(dev):
    x = rnd((8, 16, 64, 64), 2)          # attention scores shape
    a = ops.softmax_rows(H.up(x)).numpy()
    b = ops.softmax(H.up(x), 3).numpy()  # max_dim / sub / exp / sum_dim / div through tapes + reduces
    H.assert_close(a, b, 1e-5, 1e-10)
    assert np.allclose(a.sum(axis=-1), 1.0, atol=1e-5)


def test_attention_mask_rows(dev):
    x = rnd((4, 8, 32, 32), 3)
    x[:, :, :, 20:] = -1.0e9              # mask_fill(-1e9) as MHA does (mha.rs:282-285)
    got = ops.softmax_rows(H.up(x)).numpy()
    assert np.all(got[..., 20:] == 0.0)
    H.assert_close(got, oracle.softmax(x, 3), 1e-5, 1e-10)


@pytest.mark.parametrize("shape", [(64, 512), (4, 256, 512), (8, 1024), (3, 5000), (10, 36)])
def test_layer_norm_rows(dev, shape):
    x = rnd(shape, 4)
    d = shape[-1]
    g, b = rnd((d,), 5, 0.5, 1.5), rnd((d,), 6, -0.5, 0.5)
    got = ops.layer_norm(H.up(x), H.up(g), H.up(b), 1e-5).numpy()
    H.assert_close(got, oracle.layer_norm(x, g, b, 1e-5), 1e-5, 2e-6, "layer_norm")
    got = ops.layer_norm(H.up(x), None, None, 1e-5).numpy()
    H.assert_close(got, oracle.layer_norm(x, None, None, 1e-5), 1e-5, 2e-6, "layer_norm no affine")


def test_non_contiguous_last_axis_is_rejected(dev):
    from burn_b200 import _abi as abi
    t = H.up(rnd((8, 16), 7)).swap_dims(0, 1)
    with pytest.raises(abi.B200Error) as e:
        ops.softmax_rows(t)
    assert e.value.status == abi.ERR_UNSUPPORTED


# ---- backward row kernels (b200_launch_softmax_backward / b200_launch_layer_norm_backward)
BWD_SHAPES = [(7, 4), (33, 256), (2, 3, 8, 1024), (5, 2048), (3, 1000), (2, 4100), (1, 1, 6, 6)]


@pytest.mark.parametrize("shape", BWD_SHAPES)
@pytest.mark.parametrize("masked", [False, True])
def test_softmax_backward_rows(dev, shape, masked):
    """dx = (dy - sum(dy*y)) * y / div, 0 where masked — against the same formula in float64."""
    x, dy = rnd(shape, 3), rnd(shape, 4, -1.0, 1.0)
    y = oracle.softmax(x, len(shape) - 1)
    mask = None
    dmask = None
    if masked:
        mshape = (1,) * (len(shape) - 2) + tuple(shape[-2:])          # broadcast over the leading dims
        mask = np.random.default_rng(5).random(mshape) < 0.3
        dmask = H.up(mask)
    div = 8.0 if masked else 1.7
    got = ops.softmax_backward(H.up(y), H.up(dy), dmask, div).numpy()
    y64, dy64 = y.astype(np.float64), dy.astype(np.float64)
    want = (dy64 - (dy64 * y64).sum(-1, keepdims=True)) * y64 / div
    if masked:
        want = np.where(np.broadcast_to(mask, shape), 0.0, want)
        assert np.all(got[np.broadcast_to(mask, shape)] == 0.0)
    scale = np.abs(want).max() + 1e-30
    assert np.abs(got - want).max() <= 1e-5 * scale, f"max err {np.abs(got - want).max():.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("shape", [(7, 4), (64, 128), (16, 8, 512), (300, 1024), (9, 2048), (3, 1000), (2, 4100)])
@pytest.mark.parametrize("with_gamma", [True, False])
def test_layer_norm_backward_rows(dev, shape, with_gamma):
    d = shape[-1]
    x, dy = rnd(shape, 6), rnd(shape, 7, -1.0, 1.0)
    gamma = rnd((d,), 8, 0.5, 1.5) if with_gamma else None
    eps = 1e-5
    dx, dgamma, dbeta = ops.layer_norm_backward(H.up(x), H.up(dy), H.up(gamma) if with_gamma else None, eps)
    x64, dy64 = x.astype(np.float64), dy.astype(np.float64)
    mean = x64.mean(-1, keepdims=True)
    var = ((x64 - mean) ** 2).mean(-1, keepdims=True)
    rstd = 1.0 / np.sqrt(var + eps)
    xh = (x64 - mean) * rstd
    g = dy64 * (gamma.astype(np.float64) if with_gamma else 1.0)
    want_dx = (g - g.mean(-1, keepdims=True) - xh * (g * xh).mean(-1, keepdims=True)) * rstd
    want_dg = (dy64 * xh).reshape(-1, d).sum(0)
    want_db = dy64.reshape(-1, d).sum(0)
    for got, want, what in ((dx.numpy(), want_dx, "dx"), (dgamma.numpy(), want_dg, "dgamma"), (dbeta.numpy(), want_db, "dbeta")):
        scale = np.abs(want).max() + 1e-30
        err = np.abs(got - want).max()
        assert err <= 2e-5 * scale, f"{what}: max err {err:.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("shape", [(64, 256), (8, 130, 1024), (37, 3000), (5, 11)])
def test_layer_norm_backward_also_emits_the_producing_linears_bias_gradient(dev, shape):
    """b200_launch_layer_norm_backward_ex: the third partial is colsum(dx) — linear_bias_backward of the Linear that
    produced the normalised tensor (crates/burn-backend/src/backend/ops/modules/linear.rs:117-128) — and asking for it
    changes nothing else."""
    d = shape[-1]
    x, dy, gamma = rnd(shape, 16), rnd(shape, 17, -1.0, 1.0), rnd((d,), 18, 0.5, 1.5)
    dx0, dg0, db0 = ops.layer_norm_backward(H.up(x), H.up(dy), H.up(gamma), 1e-5)
    dx, dg, db, dxsum = ops.layer_norm_backward(H.up(x), H.up(dy), H.up(gamma), 1e-5, want_dx_sum=True)
    H.assert_exact(dx.numpy(), dx0.numpy())
    H.assert_exact(dg.numpy(), dg0.numpy())
    H.assert_exact(db.numpy(), db0.numpy())
    want = dx.numpy().astype(np.float64).reshape(-1, d).sum(0)
    scale = np.abs(dx.numpy()).reshape(-1, d).sum(0).max() + 1e-30
    assert np.abs(dxsum.numpy() - want).max() <= 2e-6 * scale


@pytest.mark.parametrize("n,v", [(8, 16), (33, 1001), (64, 4096), (5, 50257), (4, 50260), (450, 40004), (3, 30000)])
def test_softmax_cross_entropy_value_and_gradient(dev, n, v):
    """picked = log_softmax(x)[t] and dx = (softmax(x) − onehot(t))·scale against the oracle's log_softmax
    (the chain CrossEntropyLoss::forward_default records) — and in place over the logits."""
    from burn_b200 import _abi as abi
    x = rnd((n, v), 9, -4.0, 4.0)
    t = np.random.default_rng(10).integers(0, v, n).astype(np.int32)
    scale = 1.0 / n
    picked, dx = ops.softmax_cross_entropy(H.up(x), H.up(t), scale)
    logp = oracle.log_softmax(x, 1)
    H.assert_close(picked.numpy(), logp[np.arange(n), t], 1e-5, 1e-6, "picked")
    want = np.exp(logp.astype(np.float64))
    want[np.arange(n), t] -= 1.0
    want *= scale
    got = dx.numpy()
    assert np.abs(got - want).max() <= 1e-6 * scale + 1e-5 * np.abs(want).max()
    dxi = H.up(x)
    _, same = ops.softmax_cross_entropy(dxi, H.up(t.astype(np.int64)), scale, inplace=True)
    assert same is dxi and np.array_equal(dxi.numpy(), got)


def test_softmax_cross_entropy_many_rows_pipelined(dev):
    """More rows than CTAs: every CTA walks several rows, prefetching row r + grid into the smem slots the
    gradient pass of row r has just freed — values identical to the one-row-per-CTA case above."""
    n, v = 1000, 50260
    x = rnd((n, v), 11, -4.0, 4.0)
    t = np.random.default_rng(12).integers(0, v, n).astype(np.int32)
    picked, dx = ops.softmax_cross_entropy(H.up(x), H.up(t), 1.0 / n)
    for r in (0, 1, 147, 148, 149, 500, 999):
        logp = oracle.log_softmax(x[r:r + 1], 1)
        H.assert_close(picked.numpy()[r:r + 1], logp[0, t[r]:t[r] + 1], 1e-5, 1e-6, f"picked row {r}")
        want = np.exp(logp.astype(np.float64))
        want[0, t[r]] -= 1.0
        want /= n
        assert np.abs(dx.numpy()[r:r + 1] - want).max() <= 1e-6 / n + 1e-5 * np.abs(want).max(), f"row {r}"


def test_softmax_cross_entropy_invalid_target_is_reported(dev):
    """A target outside [0, V) makes the reference's gather panic; here the row contributes a defined 0 to
    `picked` (never uninitialised memory) and the next sync reports B200_ERR_SHAPE."""
    from burn_b200 import _abi as abi
    from burn_b200 import device as dv
    x = rnd((4, 64), 13, -1.0, 1.0)
    t = np.array([1, 64, -1, 3], dtype=np.int32)
    picked, _ = ops.softmax_cross_entropy(H.up(x), H.up(t), 0.25)
    with pytest.raises(abi.B200Error) as e:
        dv.sync()
    assert e.value.status == abi.ERR_SHAPE
    got = picked.numpy()
    assert got[1] == 0.0 and got[2] == 0.0 and got[0] != 0.0
