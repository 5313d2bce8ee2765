"""GPU: fused attention forward (b200_launch_attention, attention.cu) against the reference semantics of
attention_fallback (crates/burn-backend/src/backend/ops/modules/attention.rs:15-90) evaluated in float64:
scores·scale → bool / causal mask fill → NaN-safe softmax → · v.  Products run on tf32 tensor cores:
stated tolerance 2e-3 of max|v| on the context and 2e-3 absolute on the weights."""
import numpy as np
import pytest

from burn_b200 import ops
from burn_b200.device import DeviceTensor
from tests import helpers as H

pytestmark = pytest.mark.gpu


def reference(q, k, v, mask, scale, mask_value, causal):
    q, k, v = (t.astype(np.float64) for t in (q, k, v))
    s = np.einsum("bhqd,bhkd->bhqk", q, k) * scale
    Sq, Sk = s.shape[-2:]
    if mask is not None:
        s = np.where(np.broadcast_to(mask, s.shape), mask_value, s)
    if causal:
        cm = np.arange(Sk)[None, :] > (np.arange(Sq)[:, None] + (Sk - Sq))
        s = np.where(cm, mask_value, s)
    m = np.maximum(s.max(-1, keepdims=True), np.finfo(np.float32).min)
    e = np.exp(s - m)
    p = e / np.maximum(e.sum(-1, keepdims=True), np.finfo(np.float32).tiny)
    return np.einsum("bhqk,bhkd->bhqd", p, v), p


def rnd(shape, seed, scale=1.0):
    return (np.random.default_rng(seed).standard_normal(shape) * scale).astype(np.float32)


def check(out, w, ref_o, ref_p, v):
    tol = 2e-3 * np.abs(v).max()
    err = np.abs(out.astype(np.float64) - ref_o).max()
    assert err <= tol, f"context: max err {err:.3e} > {tol:.3e}"
    if w is not None:
        werr = np.abs(w.astype(np.float64) - ref_p).max()
        assert werr <= 2e-3, f"weights: max err {werr:.3e}"
        # rows of the returned weights sum to 1 (fully masked -inf rows are all zeros: NaN-safe softmax)
        sums = w.astype(np.float64).sum(-1)
        assert (np.abs(sums - 1.0) <= 1e-5).__or__(np.abs(sums) <= 1e-30).all(), "weights rows do not sum to 1"


@pytest.mark.parametrize("B,Hh,Sq,Sk", [(1, 2, 128, 128), (2, 3, 200, 264), (1, 1, 64, 1024), (2, 2, 384, 384), (1, 2, 1024, 1024)])
@pytest.mark.parametrize("mode", ["plain", "causal", "mask", "mask_inf"])
def test_attention_forward_matches_reference(dev, B, Hh, Sq, Sk, mode):
    q, k, v = rnd((B, Hh, Sq, 64), 1, 0.5), rnd((B, Hh, Sk, 64), 2, 0.5), rnd((B, Hh, Sk, 64), 3)
    mask = None
    causal = mode == "causal"
    mask_value = -1.0e9
    if mode.startswith("mask"):
        mask = np.random.default_rng(4).random((1, 1, Sq, Sk)) < 0.3
        if mode == "mask_inf":
            mask_value = float("-inf")
            mask[0, 0, min(5, Sq - 1), :] = True          # a fully masked row → zeros (NaN-safe softmax)
    scale = 1.0 / 8.0
    out, w = ops.attention(H.up(q), H.up(k), H.up(v), H.up(mask) if mask is not None else None, scale, mask_value,
                           causal, want_weights=True)
    ref_o, ref_p = reference(q, k, v, mask, scale, mask_value, causal)
    check(out.numpy(), w.numpy(), ref_o, ref_p, v)
    if mode == "causal":
        # skipped blocks: exact zeros above the diagonal
        wn = w.numpy()
        cm = np.arange(Sk)[None, :] > (np.arange(Sq)[:, None] + (Sk - Sq))
        assert np.all(wn[..., cm] == 0.0)
    out2 = ops.attention(H.up(q), H.up(k), H.up(v), H.up(mask) if mask is not None else None, scale, mask_value, causal)
    assert np.array_equal(out2.numpy(), out.numpy())          # the weights output does not change the context


def test_attention_on_head_views_writes_the_token_major_layout(dev):
    """q/k/v as [B,S,H,dk] projections viewed [B,H,S,dk]; the context lands in [B,S,H,dk] directly."""
    B, S, Hh, dk = 2, 256, 4, 64
    qf, kf, vf = rnd((B, S, Hh * dk), 5, 0.5), rnd((B, S, Hh * dk), 6, 0.5), rnd((B, S, Hh * dk), 7)
    heads = lambda t: H.up(t).reshape((B, S, Hh, dk)).swap_dims(1, 2)
    ctx = DeviceTensor.empty((B, S, Hh, dk))
    ops.attention(heads(qf), heads(kf), heads(vf), None, 0.125, -1.0e9, True, out=ctx.swap_dims(1, 2))
    hn = lambda t: t.reshape(B, S, Hh, dk).transpose(0, 2, 1, 3)
    ref_o, _ = reference(hn(qf), hn(kf), hn(vf), None, 0.125, -1.0e9, True)
    got = ctx.numpy().transpose(0, 2, 1, 3)
    check(got, None, ref_o, None, vf)


def test_unsupported_head_dim_is_an_error(dev):
    q = H.up(rnd((1, 1, 128, 32), 8))
    with pytest.raises(Exception):
        ops.attention(q, q, q)


@pytest.mark.parametrize("B,Hh,Sq,Sk,causal", [(1, 2, 128, 128, False), (2, 2, 200, 264, False), (1, 2, 512, 512, True), (1, 1, 1024, 1024, True)])
def test_attention_backward_dq_ds(dev, B, Hh, Sq, Sk, causal):
    """dS = P ∘ (dP − rowsum(dP ∘ P))·scale and dQ = dS·K from the saved weights, against float64."""
    q, k, v = rnd((B, Hh, Sq, 64), 11, 0.5), rnd((B, Hh, Sk, 64), 12, 0.5), rnd((B, Hh, Sk, 64), 13)
    g = rnd((B, Hh, Sq, 64), 14)
    scale = 0.125
    dq_, dk_, dv_ = H.up(q), H.up(k), H.up(v)
    out, w = ops.attention(dq_, dk_, dv_, None, scale, -1.0e9, causal, want_weights=True)
    dq, ds = ops.attention_backward(H.up(g), dk_, dv_, out, w, scale, causal)
    _, p = reference(q, k, v, None, scale, -1.0e9, causal)
    g64, k64, v64 = (t.astype(np.float64) for t in (g, k, v))
    dp = np.einsum("bhqd,bhkd->bhqk", g64, v64)
    ds_ref = p * (dp - (dp * p).sum(-1, keepdims=True)) * scale
    dq_ref = np.einsum("bhqk,bhkd->bhqd", ds_ref, k64)
    for got, want, what in ((ds.numpy(), ds_ref, "dS"), (dq.numpy(), dq_ref, "dQ")):
        tol = 3e-3 * np.abs(want).max()
        err = np.abs(got - want).max()
        assert err <= tol, f"{what}: max err {err:.3e} > {tol:.3e}"
    if causal:
        cm = np.arange(Sk)[None, :] > np.arange(Sq)[:, None] + (Sk - Sq)
        assert np.all(ds.numpy()[..., cm] == 0.0)
