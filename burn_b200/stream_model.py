"""The configs[3] TransformerEncoder forward issued the way burn-nn issues it through `Fusion<B>`: as a stream of
primitive operations, with nothing hand-fused.  The host fusion layer
(burn_b200/host/fusion.cpp) decides the kernels: Matmul blocks with bias / scale / gelu epilogues, ReduceBroadcasted
blocks for softmax and layer_norm, lone view blocks, in-place elementwise outputs, and — from the second call on —
the cached plan re-executed through a new Context without running a fuser.

Op stream per module (reference file:line):
  Linear              reshape(w → [1, d, n]), matmul, reshape(b → [1, 1, n]), add      crates/burn-nn/src/modules/linear.rs:86-104,
                                                                                         crates/burn-backend/src/backend/ops/modules/linear.rs:26-41
  MultiHeadAttention  q/k/v Linear, reshape [B,S,H,dk], swap_dims(1,2), q·kᵀ, div_scalar(√dk), softmax(3), ·v,
                      swap_dims(1,2), reshape [B,S,d], output Linear                    crates/burn-nn/src/modules/attention/mha.rs:212-311
  softmax             max_dim, sub, exp, sum_dim, div                                   crates/burn-backend/src/backend/ops/activation.rs:250-256
  gelu                div_scalar, erf, add_scalar, mul, div_scalar                      activation.rs:69-76
  layer_norm          mean_dim, sub, mul, mean_dim, add_scalar, sqrt, div, reshape γ, mul, reshape β, add
                                                                                         crates/burn-backend/src/backend/ops/modules/base.rs:846-877
  encoder layer       post-norm: x = ln1(x + mha(x)); x = ln2(x + pwff(x))              crates/burn-nn/src/modules/transformer/encoder.rs:255-290
"""
from __future__ import annotations

import math

from . import _abi as abi
from . import fusion as F


class StreamEncoder:
    """Wraps a `train.Encoder`'s parameters (non-owning) and replays its forward on a FusionStream."""

    NAMES = ("wq", "bq", "wk", "bk", "wv", "bv", "wo", "bo", "w1", "b1", "w2", "b2", "g1", "be1", "g2", "be2")

    def __init__(self, stream: F.FusionStream, encoder=None, precision: int = abi.MM_TF32):
        self.st, self.prec = stream, precision
        self.layers = []
        for l in (encoder.layers if encoder is not None else ()):
            w = {name: stream.wrap(getattr(l, name).v.desc()) for name in self.NAMES}
            w["h"] = l.h
            self.layers.append(w)

    @classmethod
    def placeholders(cls, stream: F.FusionStream, d_model: int, d_ff: int, n_heads: int, n_layers: int) -> "StreamEncoder":
        """Parameter placeholders on a plan-only stream (fusion decisions without a device)."""
        self = cls(stream)
        shapes = {"wq": (d_model, d_model), "wk": (d_model, d_model), "wv": (d_model, d_model), "wo": (d_model, d_model),
                  "w1": (d_model, d_ff), "w2": (d_ff, d_model), "b1": (d_ff,)}
        for _ in range(n_layers):
            w = {name: stream.placeholder(shapes.get(name, (d_model,))) for name in cls.NAMES}
            w["h"] = n_heads
            self.layers.append(w)
        return self

    # ---- modules as op streams
    def linear(self, x: F.LazyTensor, w: F.LazyTensor, b: F.LazyTensor) -> F.LazyTensor:
        rank = len(x.shape)
        d_in, d_out = w.shape
        w3 = w.reshape((1,) * (rank - 2) + (d_in, d_out))
        y = x.matmul(w3, self.prec)
        w3.drop()
        b3 = b.reshape((1,) * (rank - 1) + (d_out,))
        z = y.add(b3)
        y.drop()
        b3.drop()
        return z

    def mha(self, x: F.LazyTensor, p: dict) -> F.LazyTensor:
        B, S, d = x.shape
        H = p["h"]
        dk = d // H

        def heads(t):
            r = t.reshape((B, S, H, dk))
            t.drop()
            v = r.swap_dims(1, 2)
            r.drop()
            return v
        q = heads(self.linear(x, p["wq"], p["bq"]))
        k = heads(self.linear(x, p["wk"], p["bk"]))
        v = heads(self.linear(x, p["wv"], p["bv"]))
        kt = k.swap_dims(2, 3)
        k.drop()
        s = q.matmul(kt, self.prec)
        q.drop()
        kt.drop()
        sc = s.div_scalar(math.sqrt(dk))
        s.drop()
        wts = F.softmax(sc, 3)
        sc.drop()
        ctx = wts.matmul(v, self.prec)
        wts.drop()
        v.drop()
        ct = ctx.swap_dims(1, 2)
        ctx.drop()
        flat = ct.reshape((B, S, d))
        ct.drop()
        out = self.linear(flat, p["wo"], p["bo"])
        flat.drop()
        return out

    def layer_norm(self, x: F.LazyTensor, gamma: F.LazyTensor, beta: F.LazyTensor) -> F.LazyTensor:
        rank = len(x.shape)
        g = gamma.reshape((1,) * (rank - 1) + (gamma.shape[0],))
        b = beta.reshape((1,) * (rank - 1) + (beta.shape[0],))
        y = F.layer_norm(x, g, b, 1e-5)
        g.drop()
        b.drop()
        return y

    def layer(self, x: F.LazyTensor, p: dict, drop_input: bool) -> F.LazyTensor:
        a = self.mha(x, p)
        r = x.add(a)
        a.drop()
        if drop_input:
            x.drop()
        x1 = self.layer_norm(r, p["g1"], p["be1"])
        r.drop()
        h = self.linear(x1, p["w1"], p["b1"])
        g = F.gelu(h)
        h.drop()
        f = self.linear(g, p["w2"], p["b2"])
        g.drop()
        r2 = x1.add(f)
        f.drop()
        x1.drop()
        y = self.layer_norm(r2, p["g2"], p["be2"])
        r2.drop()
        return y

    def forward(self, x: F.LazyTensor) -> F.LazyTensor:
        cur = x
        for i, p in enumerate(self.layers):
            cur = self.layer(cur, p, drop_input=i > 0)
        return cur
