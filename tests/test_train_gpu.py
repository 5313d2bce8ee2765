"""GPU: the transformer training-step workload (burn_b200/train.py — the stand-in for burn-nn +
burn-autodiff + burn-optim on top of our kernels) against an independent float64 PyTorch-CPU
autograd implementation of the same model with the same weights.  torch is only the cross-check
here ("torch (CPU) is available as an independent cross-check, never as the oracle").
GEMMs run in F32X3 so the comparison is tight; loss to 1e-5, gradients to 2e-4 relative."""
import math

import numpy as np
import pytest
import torch

from burn_b200 import _abi as abi
from burn_b200 import train as T
from burn_b200.device import DeviceTensor
from tests import helpers as H

pytestmark = pytest.mark.gpu


def t64(p):
    return torch.tensor(p.v.numpy().astype(np.float64), requires_grad=True)


def torch_layer(x, P, h, mask):
    B, S, d = x.shape
    dk = d // h
    lin = lambda t, w, b: t @ w + b
    q, k, v = lin(x, P["wq"], P["bq"]), lin(x, P["wk"], P["bk"]), lin(x, P["wv"], P["bv"])
    hd = lambda t: t.reshape(B, S, h, dk).transpose(1, 2)
    sc = hd(q) @ hd(k).transpose(2, 3) / math.sqrt(dk)
    if mask is not None:
        sc = sc.masked_fill(mask, -1.0e9)
    w = torch.softmax(sc, dim=-1)
    ctx = (w @ hd(v)).transpose(1, 2).reshape(B, S, d)
    x = x + lin(ctx, P["wo"], P["bo"])
    x = torch.nn.functional.layer_norm(x, (d,), P["g1"], P["be1"], 1e-5)
    hdn = torch.nn.functional.gelu(lin(x, P["w1"], P["b1"]))
    x = x + lin(hdn, P["w2"], P["b2"])
    return torch.nn.functional.layer_norm(x, (d,), P["g2"], P["be2"], 1e-5)


NAMES = ["wq", "bq", "wk", "bk", "wv", "bv", "wo", "bo", "w1", "b1", "w2", "b2", "g1", "be1", "g2", "be2"]


def compare_grads(params, tparams, rel=2e-4, abs_=2e-6):
    for p, tp in zip(params, tparams):
        assert p.g is not None, f"{p.name} has no gradient"
        got = p.g.numpy().astype(np.float64)
        want = tp.grad.numpy()
        scale = np.abs(want).max() + 1e-12
        err = np.abs(got - want).max()
        assert err <= rel * scale + abs_, f"{p.name}: max grad error {err:.3e} vs scale {scale:.3e}"


def test_encoder_forward_backward_matches_float64_autograd(dev):
    d, ff, h, L, B, S = 32, 64, 4, 2, 2, 8
    enc = T.Encoder(0, d, ff, h, L)
    x = np.random.default_rng(1).standard_normal((B, S, d)).astype(np.float32)
    tape = T.Tape(abi.MM_F32X3)
    xin = T.Var(H.up(x), True, "x")
    loss = T.mean_square(tape, enc.forward(tape, xin))
    tape.backward()
    # reference
    tx = torch.tensor(x.astype(np.float64), requires_grad=True)
    tps = []
    y = tx
    for layer in enc.layers:
        P = {n: t64(p) for n, p in zip(NAMES, layer.params())}
        tps += [P[n] for n in NAMES]
        y = torch_layer(y, P, h, None)
    tl = (y * y).mean()
    tl.backward()
    assert abs(float(loss.v.numpy()[0]) - float(tl)) <= 1e-5 * abs(float(tl)) + 1e-7
    compare_grads(enc.params(), tps)
    got = xin.g.numpy().astype(np.float64)
    assert np.abs(got - tx.grad.numpy()).max() <= 2e-4 * np.abs(tx.grad.numpy()).max() + 2e-6


def test_language_model_loss_and_grads_with_causal_mask(dev):
    vocab, S, d, ff, h, L, B = 53, 8, 32, 64, 4, 1, 3   # odd vocab: exercises the unfused-bias path
    lm = T.LanguageModel(3, vocab, S, d, ff, h, L)
    rng = np.random.default_rng(4)
    tok = rng.integers(0, vocab, (B, S)).astype(np.int32)
    tgt = rng.integers(0, vocab, (B, S)).astype(np.int32)
    pos = np.tile(np.arange(S, dtype=np.int32), (B, 1))
    causal = np.triu(np.ones((S, S), dtype=bool), k=1)[None, None]       # true = masked
    tape = T.Tape(abi.MM_F32X3)
    loss = lm.loss(tape, H.up(tok), H.up(tgt), H.up(pos), H.up(causal))
    tape.backward()
    # reference
    tt, tp = t64(lm.tok), t64(lm.pos)
    x = (tt[torch.tensor(tok, dtype=torch.long)] + tp[torch.tensor(pos, dtype=torch.long)]) / 2
    P = {n: t64(p) for n, p in zip(NAMES, lm.enc.layers[0].params())}
    hdn = torch_layer(x, P, h, torch.tensor(causal))
    two, tbo = t64(lm.wout), t64(lm.bout)
    logits = (hdn @ two + tbo).reshape(B * S, vocab)
    tl = torch.nn.functional.cross_entropy(logits, torch.tensor(tgt.reshape(-1), dtype=torch.long))
    tl.backward()
    assert abs(float(loss.v.numpy()[0]) - float(tl)) <= 1e-5 * abs(float(tl))
    compare_grads(lm.params(), [tt, tp] + [P[n] for n in NAMES] + [two, tbo])


def test_adam_step_matches_formula(dev):
    """AdaptiveMomentum::transform + Adam::step (adam.rs:149-210, :80-84) in float64."""
    rng = np.random.default_rng(5)
    w, g = rng.standard_normal((33, 20)).astype(np.float32), rng.standard_normal((33, 20)).astype(np.float32)
    p = T.Param(w, "w")
    opt = T.Adam(lr=1e-2, beta1=0.9, beta2=0.999, eps=1e-5)
    m = np.zeros_like(w, dtype=np.float64)
    s = np.zeros_like(w, dtype=np.float64)
    ref = w.astype(np.float64)
    for t in range(1, 4):
        p.g = H.up(g)
        opt.step([p])
        m = 0.9 * m + 0.1 * g
        s = 0.999 * s + 0.001 * g.astype(np.float64) ** 2
        bc2s = np.sqrt(1 - 0.999 ** t)
        ref = ref - 1e-2 * (m * (bc2s / (1 - 0.9 ** t))) / (np.sqrt(s) + 1e-5 * bc2s)
    assert np.allclose(p.v.numpy(), ref, rtol=2e-5, atol=1e-6)


def test_tf32_and_bf16_training_losses_stay_close_to_f32x3(dev):
    """Stated tolerance for reduced-precision GEMMs on the end-to-end loss: 2e-3 (tf32), 2e-2 (bf16)."""
    d, ff, h, L, B, S = 64, 128, 4, 2, 4, 16
    x = np.random.default_rng(7).standard_normal((B, S, d)).astype(np.float32)
    losses = {}
    for prec in (abi.MM_F32X3, abi.MM_TF32, abi.MM_BF16):
        enc = T.Encoder(6, d, ff, h, L)
        tape = T.Tape(prec)
        losses[prec] = float(T.mean_square(tape, enc.forward(tape, T.Var(H.up(x), False))).v.numpy()[0])
    ref = losses[abi.MM_F32X3]
    assert abs(losses[abi.MM_TF32] - ref) <= 2e-3 * abs(ref)
    assert abs(losses[abi.MM_BF16] - ref) <= 2e-2 * abs(ref)


def test_param_arena_matches_per_parameter_adam_bit_for_bit(dev):
    """Flat buckets (direct-to-bucket gradients + multi-tensor Adam) change where tensors live,
    not a single rounding: parameters after two steps equal the per-parameter path exactly."""
    d, ff, h, L, B, S = 32, 64, 4, 1, 2, 8
    x = np.random.default_rng(9).standard_normal((B, S, d)).astype(np.float32)
    results = []
    for use_arena in (False, True):
        enc = T.Encoder(12, d, ff, h, L)
        params = enc.params()
        opt = T.Adam(lr=1e-2)
        arena = T.ParamArena(params, None, bucket_bytes=8 << 10) if use_arena else None
        for _ in range(2):
            tape = T.Tape(abi.MM_F32X3)
            T.mean_square(tape, enc.forward(tape, T.Var(H.up(x), False)))
            tape.backward()
            opt.advance()
            if use_arena:
                arena.wait()
                opt.apply_arena(arena)
            else:
                opt.apply(params)
            T.Adam.zero_grad(params)
        results.append([p.v.numpy() for p in params])
        if use_arena:
            assert len(arena.buckets) > 1
    for a, b in zip(*results):
        assert np.array_equal(a, b)


def test_fused_attention_training_step_matches_the_unfused_chain(dev):
    """dk = 64, tf32: the flash kernels (statistics saved, weights recomputed in the backward) and the
    weights-saving fused kernels against the scores-GEMM → softmax → context-GEMM chain — same loss and gradients
    to tf32 accuracy, with the causal flag (block skipping)."""
    vocab, S, d, ff, h, L, B = 64, 256, 128, 256, 2, 1, 2
    rng = np.random.default_rng(21)
    tok = rng.integers(0, vocab, (B, S)).astype(np.int32)
    tgt = rng.integers(0, vocab, (B, S)).astype(np.int32)
    pos = np.tile(np.arange(S, dtype=np.int32), (B, 1))
    causal = np.triu(np.ones((S, S), dtype=bool), k=1)[None, None]
    results = {}
    for fused in ("chain", "flash", "weights"):
        T._ATTN_MODE = fused
        lm = T.LanguageModel(3, vocab, S, d, ff, h, L)
        tape = T.Tape(abi.MM_TF32)
        loss = lm.loss(tape, H.up(tok), H.up(tgt), H.up(pos), H.up(causal))
        tape.backward()
        results[fused] = (float(loss.v.numpy()[0]), [p.g.numpy() for p in lm.params()])
    T._ATTN_MODE = "flash"
    for mode in ("flash", "weights"):
        assert abs(results[mode][0] - results["chain"][0]) <= 2e-3 * abs(results["chain"][0])
        for a, b in zip(results[mode][1], results["chain"][1]):
            scale = np.abs(b).max()
            # 1e-6 floor: the key-bias gradient is mathematically zero (softmax ignores a shift of all keys)
            assert np.abs(a - b).max() <= 2e-2 * scale + 1e-6, mode
