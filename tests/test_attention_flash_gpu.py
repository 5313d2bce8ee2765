"""GPU: flash-style attention (b200_launch_attention_flash / _backward, attention_flash.cu) against the
reference semantics of attention_fallback (crates/burn-backend/src/backend/ops/modules/attention.rs:15-90) and
of burn-autodiff's reverse walk over it, evaluated in float64.  No [B,H,Sq,Sk] tensor exists on the device side:
the forward saves per-row statistics, the backward recomputes the weights.  tf32 tensor-core products: stated
tolerance 2e-3 of max|v| on the context, 3e-3 of max|grad| on dq / dk / dv."""
import numpy as np
import pytest

from burn_b200 import ops
from burn_b200.device import DeviceTensor
from tests import helpers as H
from tests.test_attention_gpu import reference, rnd

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(120)]

SHAPES = [(1, 2, 128, 128), (2, 3, 200, 264), (1, 1, 64, 1024), (2, 2, 384, 384), (1, 2, 1024, 1024), (1, 1, 320, 192)]


def make_mask(mode, Sq, Sk):
    mask, mask_value = None, -1.0e9
    if mode.startswith("mask"):
        mask = np.random.default_rng(4).random((1, 1, Sq, Sk)) < 0.3
        if mode == "mask_inf":
            mask_value = float("-inf")
            mask[0, 0, min(5, Sq - 1), :] = True          # a fully masked row -> zeros (NaN-safe softmax)
        if mode == "mask_row":
            mask[0, 0, min(7, Sq - 1), :] = True          # fully masked with -1e9 -> uniform weights, zero gradient to q/k
    return mask, mask_value


@pytest.mark.parametrize("B,Hh,Sq,Sk", SHAPES)
@pytest.mark.parametrize("mode", ["plain", "causal", "mask", "mask_inf", "mask_row"])
def test_flash_forward_matches_reference(dev, B, Hh, Sq, Sk, mode):
    q, k, v = rnd((B, Hh, Sq, 64), 1, 0.5), rnd((B, Hh, Sk, 64), 2, 0.5), rnd((B, Hh, Sk, 64), 3)
    mask, mask_value = make_mask(mode, Sq, Sk)
    causal = mode == "causal"
    out, stats = ops.attention_flash(H.up(q), H.up(k), H.up(v), H.up(mask) if mask is not None else None, 0.125,
                                     mask_value, causal)
    ref_o, ref_p = reference(q, k, v, mask, 0.125, mask_value, causal)
    tol = 2e-3 * np.abs(v).max()
    err = np.abs(out.numpy().astype(np.float64) - ref_o).max()
    assert err <= tol, f"context: max err {err:.3e} > {tol:.3e}"
    # the saved statistics reproduce the weights: 2^(s2 - m2) / l
    st = stats.numpy().astype(np.float64)
    s2 = np.einsum("bhqd,bhkd->bhqk", q.astype(np.float64), k.astype(np.float64)) * 0.125 * np.log2(np.e)
    # a filled score is the f32 product mask_value * log2(e) in the kernel; at -1e9 the f32 spacing is 128, so the
    # float64 value would be off by a whole power of two on a fully masked row
    with np.errstate(invalid="ignore"):
        mask2 = float(np.float32(mask_value) * np.float32(np.log2(np.e)))
    if mask is not None:
        s2 = np.where(np.broadcast_to(mask, s2.shape), mask2, s2)
    if causal:
        cm = np.arange(Sk)[None, :] > (np.arange(Sq)[:, None] + (Sk - Sq))
        s2 = np.where(cm, mask2, s2)
    with np.errstate(over="ignore", invalid="ignore"):
        p = np.exp2(s2 - st[..., 0:1]) * st[..., 1:2]
    assert np.abs(p - ref_p).max() <= 2e-3


def test_flash_forward_equals_the_weights_kernel(dev):
    """Same tf32 QK^T products and base-2 softmax; the flash kernel rounds the UNNORMALISED weights to tf32 for the
    PV product and divides afterwards, the weights kernel normalises first — so they agree to tf32 accuracy of P
    (2^-11 relative per term), not bit for bit: 5e-4 of max|v|."""
    q, k, v = rnd((2, 4, 512, 64), 21, 0.5), rnd((2, 4, 512, 64), 22, 0.5), rnd((2, 4, 512, 64), 23)
    a = ops.attention(H.up(q), H.up(k), H.up(v), None, 0.125, -1.0e9, True).numpy()
    b, _ = ops.attention_flash(H.up(q), H.up(k), H.up(v), None, 0.125, -1.0e9, True)
    assert np.abs(a - b.numpy()).max() <= 5e-4 * np.abs(v).max()


def grads_reference(q, k, v, g, mask, scale, mask_value, causal):
    _, p = reference(q, k, v, mask, scale, mask_value, causal)
    g64, q64, k64, v64 = (t.astype(np.float64) for t in (g, q, k, v))
    dp = np.einsum("bhqd,bhkd->bhqk", g64, v64)
    ds = p * (dp - (dp * p).sum(-1, keepdims=True)) * scale
    Sq, Sk = ds.shape[-2:]
    if mask is not None:                                  # mask_fill backward: no gradient through a filled score
        ds = np.where(np.broadcast_to(mask, ds.shape), 0.0, ds)
    if causal:
        cm = np.arange(Sk)[None, :] > (np.arange(Sq)[:, None] + (Sk - Sq))
        ds = np.where(cm, 0.0, ds)
    return (np.einsum("bhqk,bhkd->bhqd", ds, k64), np.einsum("bhqk,bhqd->bhkd", ds, q64),
            np.einsum("bhqk,bhqd->bhkd", p, g64))


@pytest.mark.parametrize("B,Hh,Sq,Sk", SHAPES)
@pytest.mark.parametrize("mode", ["plain", "causal", "mask", "mask_row"])
def test_flash_backward_matches_reference(dev, B, Hh, Sq, Sk, mode):
    q, k, v = rnd((B, Hh, Sq, 64), 11, 0.5), rnd((B, Hh, Sk, 64), 12, 0.5), rnd((B, Hh, Sk, 64), 13)
    g = rnd((B, Hh, Sq, 64), 14)
    mask, mask_value = make_mask(mode, Sq, Sk)
    causal = mode == "causal"
    dq_, dk_, dv_ = H.up(q), H.up(k), H.up(v)
    dm = H.up(mask) if mask is not None else None
    out, stats = ops.attention_flash(dq_, dk_, dv_, dm, 0.125, mask_value, causal)
    dq, dk, dv = ops.attention_flash_backward(H.up(g), dq_, dk_, dv_, out, stats, dm, 0.125, mask_value, causal)
    want = grads_reference(q, k, v, g, mask, 0.125, mask_value, causal)
    for got, ref, what in ((dq.numpy(), want[0], "dQ"), (dk.numpy(), want[1], "dK"), (dv.numpy(), want[2], "dV")):
        tol = 3e-3 * np.abs(ref).max()
        err = np.abs(got - ref).max()
        assert err <= tol, f"{what}: max err {err:.3e} > {tol:.3e}"


def test_flash_on_head_views_and_determinism(dev):
    """[B,S,H,dk] projections viewed [B,H,S,dk] in and out (the layout burn-nn's MHA produces), and two runs of
    the backward are bit-identical (no atomics anywhere)."""
    B, S, Hh, dk = 2, 256, 4, 64
    qf, kf, vf, gf = rnd((B, S, Hh * dk), 5, 0.5), rnd((B, S, Hh * dk), 6, 0.5), rnd((B, S, Hh * dk), 7), rnd((B, S, Hh * dk), 8)
    heads = lambda t: H.up(t).reshape((B, S, Hh, dk)).swap_dims(1, 2)
    q, k, v, g = heads(qf), heads(kf), heads(vf), heads(gf)
    ctx = DeviceTensor.empty((B, S, Hh, dk))
    _, stats = ops.attention_flash(q, k, v, None, 0.125, -1.0e9, True, out=ctx.swap_dims(1, 2))
    hn = lambda t: t.reshape(B, S, Hh, dk).transpose(0, 2, 1, 3)
    ref_o, _ = reference(hn(qf), hn(kf), hn(vf), None, 0.125, -1.0e9, True)
    assert np.abs(ctx.numpy().transpose(0, 2, 1, 3) - ref_o).max() <= 2e-3 * np.abs(vf).max()
    bufs = [DeviceTensor.empty((B, S, Hh, dk)) for _ in range(6)]
    runs = []
    for i in range(2):
        d = [b.swap_dims(1, 2) for b in bufs[3 * i:3 * i + 3]]
        ops.attention_flash_backward(g, q, k, v, ctx.swap_dims(1, 2), stats, None, 0.125, -1.0e9, True, *d)
        runs.append([b.numpy() for b in bufs[3 * i:3 * i + 3]])
    want = grads_reference(hn(qf), hn(kf), hn(vf), hn(gf), None, 0.125, -1.0e9, True)
    for a, b, ref in zip(runs[0], runs[1], want):
        assert np.array_equal(a, b)
        assert np.abs(a.transpose(0, 2, 1, 3) - ref).max() <= 3e-3 * np.abs(ref).max()


def test_flash_full_config4_shape(dev):
    """configs[4] attention shape [8, 16, 1024, 64] causal: forward and backward against float64 on two heads
    (the float64 reference of all 128 heads would take minutes on the host)."""
    B, Hh, S = 8, 16, 1024
    q, k, v, g = rnd((B, Hh, S, 64), 31, 0.5), rnd((B, Hh, S, 64), 32, 0.5), rnd((B, Hh, S, 64), 33), rnd((B, Hh, S, 64), 34)
    dq_, dk_, dv_ = H.up(q), H.up(k), H.up(v)
    out, stats = ops.attention_flash(dq_, dk_, dv_, None, 0.125, -1.0e9, True)
    dq, dk, dv = ops.attention_flash_backward(H.up(g), dq_, dk_, dv_, out, stats, None, 0.125, -1.0e9, True)
    o, gq, gk, gv = out.numpy(), dq.numpy(), dk.numpy(), dv.numpy()
    for (b, h) in ((0, 0), (7, 15)):
        sl = (slice(b, b + 1), slice(h, h + 1))
        ref_o, _ = reference(q[sl], k[sl], v[sl], None, 0.125, -1.0e9, True)
        assert np.abs(o[sl] - ref_o).max() <= 2e-3 * np.abs(v).max()
        want = grads_reference(q[sl], k[sl], v[sl], g[sl], None, 0.125, -1.0e9, True)
        for got, ref in zip((gq[sl], gk[sl], gv[sl]), want):
            assert np.abs(got - ref).max() <= 3e-3 * np.abs(ref).max()


def test_per_item_backward_kernels_stay_covered(dev):
    """The host picks the persistent dQ / dK/dV kernels for causal masks and short loops; B200_FA_PERSIST=0 (read once
    per process) forces the one-CTA-per-item kernels for every shape.  Run the causal backward cases under it in a child
    process so that both kernel families are held to the float64 reference."""
    import os, subprocess, sys
    env = dict(os.environ, B200_FA_PERSIST="0")
    r = subprocess.run([sys.executable, "-m", "pytest", __file__, "-m", "gpu", "-q", "-x", "--timeout", "100",
                        "-k", "test_flash_backward_matches_reference and causal", "-p", "no:cacheprovider"],
                       env=env, capture_output=True, text=True, timeout=300,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert " passed" in r.stdout
