"""Training-step workload generator for BASELINE configs 4/5 (transformer encoder / decoder LM).

NOT part of the kernel product: this module plays the role of the reference's *callers* of the
hot path — burn-nn modules, burn-autodiff's reverse tape and burn-optim's Adam — so that the
fused elementwise / reduce / matmul kernels and the NCCL all-reduce can be measured on the op
streams those crates issue.  It mirrors, op for op:
  Linear                 crates/burn-backend/src/backend/ops/modules/linear.rs:17-128
  LayerNorm              crates/burn-backend/src/backend/ops/modules/base.rs:846-877
  MultiHeadAttention     crates/burn-nn/src/modules/attention/mha.rs:212-311
  PositionWiseFeedForward crates/burn-nn/src/modules/transformer/pwff.rs:105-111
  TransformerEncoderLayer crates/burn-nn/src/modules/transformer/encoder.rs:245-303 (post-norm default)
  matmul backward        crates/burn-autodiff/src/ops/tensor.rs:560-616 (grad·rhsᵀ, lhsᵀ·grad)
  cross-entropy          crates/burn-nn/src/loss/cross_entropy.rs:171-197 (log_softmax, gather, mean)
  Adam                   crates/burn-optim/src/optim/adam.rs:149-210
All tensor math runs in libburn_b200.so through burn_b200.ops; this file only sequences launches.
"""
from __future__ import annotations

import ctypes as C
import os
import math
from typing import Callable, Sequence

import numpy as np

from . import _abi as abi
from . import device as dv
from . import ops
from .device import DeviceTensor, TapeBuilder


# ------------------------------------------------------------------ reverse tape (burn-autodiff)
class Var:
    """A tensor on the autodiff tape."""
    __slots__ = ("v", "g", "requires_grad", "name", "aux")

    def __init__(self, v: DeviceTensor, requires_grad: bool = False, name: str = ""):
        self.v, self.g, self.requires_grad, self.name = v, None, requires_grad, name
        self.aux = None     # notes between backward steps (e.g. a bias gradient the consumer's backward already produced)


class Tape:
    def __init__(self, precision: int = abi.MM_TF32):
        self.steps: list[Callable[[], None]] = []
        self.precision = precision

    def add(self, fn: Callable[[], None]) -> None:
        self.steps.append(fn)

    def backward(self) -> None:
        for fn in reversed(self.steps):
            fn()
        self.steps.clear()


def accumulate(var: Var, g: DeviceTensor) -> None:
    if not var.requires_grad:
        return
    var.g = g if var.g is None else ops.float_add(var.g, g)
    hook = getattr(var, "on_grad", None)
    if hook is not None:
        hook(var)


def _run(tb: TapeBuilder, inputs, shape, n_out=1):
    outs = [DeviceTensor.empty(shape) for _ in range(n_out)]
    dv.launch_elemwise(tb.build(), inputs, outs, shape)
    return outs[0] if n_out == 1 else outs


# ------------------------------------------------------------------ ops with backward
def matmul(tape: Tape, a: Var, b: Var, out: DeviceTensor | None = None, epilogue=None, epi_inputs=()) -> Var:
    y = Var(_mm(a.v, b.v, tape.precision, out, epilogue, epi_inputs), a.requires_grad or b.requires_grad)

    def bw():
        if y.g is None:
            return
        nd = a.v.ndim
        if a.requires_grad:
            accumulate(a, _reduce_to(_mm(y.g, b.v.swap_dims(nd - 2, nd - 1), tape.precision), a.v.shape))
        if b.requires_grad:
            accumulate(b, _reduce_to(_mm(a.v.swap_dims(nd - 2, nd - 1), y.g, tape.precision), b.v.shape))
    tape.add(bw)
    return y


def _mm(a: DeviceTensor, b: DeviceTensor, precision, out=None, epilogue=None, epi_inputs=()):
    if out is None:
        return ops.float_matmul(a, b, precision, epilogue, epi_inputs)
    ad, bd, cd = a.desc(), b.desc(), out.desc()
    lib = abi.load()
    wsb = C.c_uint64()
    abi.check(lib.b200_matmul_workspace_bytes(C.byref(ad), C.byref(bd), precision, C.byref(wsb)))
    ws = dv.Storage(wsb.value) if wsb.value else None
    epi, n_epi = dv._descs(epi_inputs)
    abi.check(lib.b200_launch_matmul(C.byref(ad), C.byref(bd), C.byref(cd), precision,
                                     C.byref(epilogue) if epilogue is not None else None, epi, n_epi,
                                     ws.ptr if ws else None, wsb.value, None))
    return out


def _reduce_to(g: DeviceTensor, shape) -> DeviceTensor:
    """Sums broadcast batch dims back to `shape` (autodiff's broadcast backward = SumDim chains)."""
    for d, (gs, s) in enumerate(zip(g.shape, shape)):
        if gs != s:
            g = ops.float_sum_dim(g, d)
    return g


def linear(tape: Tape, x: Var, w: Var, b: Var | None, residual: Var | None = None) -> Var:
    """x[..., d_in] · W[d_in, d_out] + b [+ residual], batch folded into M (linear.rs:26-41).  The bias add —
    and the residual add of the encoder layer that follows it (encoder.rs:263-264,283) — are the GEMM's
    fuse-on-write epilogue (MatmulOptimization): same two roundings, no extra pass over the activations."""
    lead = x.v.shape[:-1]
    m = int(np.prod(lead))
    n_out = w.v.shape[1]
    x2 = x.v.reshape((m, x.v.shape[-1]))
    epi = epi_in = None
    fuse = b is not None and n_out % 4 == 0     # the fused epilogue needs N % 4 == 0
    if fuse:
        tb = TapeBuilder().op("ADD_F", ("in", 0), ("in", 1), out=None if residual is not None else 0)
        epi_in = [b.v.reshape((1, n_out))]
        if residual is not None:
            tb.op("ADD_F", ("in", 2), "acc", out=0)                 # x + (acc + bias), operand order of `x + y`
            epi_in.append(residual.v.reshape((m, n_out)))
        epi = tb.build()
    y2 = ops.float_matmul(x2, w.v, tape.precision, epi, epi_in or ())
    if not fuse:
        if b is not None:
            y2 = ops.float_add(y2, b.v.reshape((1, n_out)))
        if residual is not None:
            y2 = ops.float_add(residual.v.reshape((m, n_out)), y2)
    y = Var(y2.reshape(tuple(lead) + (n_out,)), True)
    if b is not None and b.requires_grad:
        y.aux = {"wants_bias_grad": True}     # a consumer whose backward streams dy anyway may leave colsum(dy) here

    def bw():
        if y.g is None:
            return
        if residual is not None:
            accumulate(residual, y.g)
        g2 = y.g.reshape((m, w.v.shape[1]))
        if x.requires_grad:
            # (folding `x.g + dx` into this GEMM's epilogue was measured: the full-size epilogue operand costs the GEMM
            # +50 us against a 20 us add — the separate fused add stays)
            accumulate(x, ops.float_matmul(g2, w.v.swap_dims(0, 1), tape.precision).reshape(x.v.shape))
        if w.requires_grad:
            slot = getattr(w, "grad_slot", None)
            if slot is not None and w.g is None:      # the GEMM writes the bucket slot directly
                accumulate(w, _mm(x2.swap_dims(0, 1), g2, tape.precision, out=slot))
            else:
                accumulate(w, ops.float_matmul(x2.swap_dims(0, 1), g2, tape.precision))
        if b is not None and b.requires_grad:
            ready = (y.aux or {}).get("bias_grad")
            if ready is not None and ready[1] is y.g:                    # colsum of exactly this gradient, made by its producer
                accumulate(b, ready[0].reshape(b.v.shape))
            else:
                accumulate(b, ops.float_sum_dim(g2, 0).reshape(b.v.shape))   # linear_bias_backward: column reduce
    tape.add(bw)
    return y


def add(tape: Tape, a: Var, b: Var) -> Var:
    y = Var(ops.float_add(a.v, b.v), True)

    def bw():
        if y.g is not None:
            accumulate(a, y.g)
            accumulate(b, y.g)
    tape.add(bw)
    return y


GELU_BWD_EXACT = os.environ.get("B200_GELU_BWD", "reference") == "exact"
_SYNC_DEFER = os.environ.get("B200_SYNC_DEFER", "0") == "1"


def gelu_backward_tape() -> TapeBuilder:
    """B::gelu_backward as burn-autodiff calls it (crates/burn-autodiff/src/ops/activation.rs:35): the trait
    default is the derivative of the TANH approximation, not of the erf form the forward uses
    (crates/burn-backend/src/backend/ops/activation.rs:98-128) — same 20 primitive ops, same operation order,
    one fused tape over (x, dy).  B200_GELU_BWD=exact selects the analytic erf derivative instead (a deliberate
    deviation from the reference trajectory, kept for comparison)."""
    tb = TapeBuilder()
    if GELU_BWD_EXACT:
        # d/dx [x(1+erf(x/sqrt2))/2] = (1+erf(x/sqrt2))/2 + x*exp(-x^2/2)/sqrt(2*pi)
        tb.op("MUL_F", ("in", 0), ("in", 0))
        tb.op("MUL_F", "acc", ("f", -0.5))
        tb.op("EXP_F", "acc")
        tb.op("MUL_F", "acc", ("in", 0))
        tb.op("MUL_F", "acc", ("f", 0.3989422804014327), tmp=0)
        tb.op("DIV_F", ("in", 0), ("f", 1.4142135623730951))
        tb.op("ERF_F", "acc")
        tb.op("ADD_F", "acc", ("f", 1.0))
        tb.op("MUL_F", "acc", ("f", 0.5))
        tb.op("ADD_F", "acc", ("tmp", 0))
        tb.op("MUL_F", "acc", ("in", 1), out=0)
        return tb
    tb.op("POW_F", ("in", 0), ("f", 3.0), tmp=0)                 # x3 = powi_scalar(x, 3) -> powf_scalar_impl
    tb.op("MUL_F", ("in", 0), ("f", 0.797885), tmp=1)            # c2
    tb.op("MUL_F", ("tmp", 0), ("f", 0.0356774))                 # c1
    tb.op("ADD_F", "acc", ("tmp", 1))                            # inner1 = c1 + c2
    tb.op("TANH_F", "acc", tmp=1)                                # tanh
    tb.op("MUL_F", ("in", 0), ("f", 0.398942), tmp=2)            # c4
    tb.op("MUL_F", ("tmp", 0), ("f", 0.0535161))                 # c3
    tb.op("ADD_F", "acc", ("tmp", 2), tmp=0)                     # inner2 = c3 + c4
    tb.op("MUL_F", ("tmp", 1), ("tmp", 1))                       # powi_scalar(tanh, 2) = tanh * tanh
    tb.op("NEG_F", "acc")
    tb.op("ADD_F", "acc", ("f", 1.0))                            # sech = 1 - tanh^2
    tb.op("MUL_F", ("tmp", 0), "acc")                            # inner2 * sech
    tb.op("ADD_F", "acc", ("f", 0.5), tmp=0)                     # y2
    tb.op("MUL_F", ("tmp", 1), ("f", 0.5))                       # y1
    tb.op("ADD_F", "acc", ("tmp", 0))                            # y = y1 + y2
    tb.op("MUL_F", "acc", ("in", 1), out=0)                      # y * grad
    return tb


def gelu(tape: Tape, x: Var) -> Var:
    y = Var(ops.gelu(x.v), True)

    def bw():
        if y.g is None or not x.requires_grad:
            return
        accumulate(x, _run(gelu_backward_tape(), [x.v, y.g], x.v.shape))
    tape.add(bw)
    return y


def layer_norm(tape: Tape, x: Var, gamma: Var, beta: Var, eps: float = 1e-5) -> Var:
    y = Var(ops.layer_norm(x.v, gamma.v, beta.v, eps), True)

    def bw():
        if y.g is None:
            return
        # one row-resident kernel: statistics recomputed on chip, dx written once, dgamma/dbeta as
        # per-CTA partials finished by a column reduce
        want = bool(x.aux and x.aux.get("wants_bias_grad")) and x.g is None
        if want:       # x = Linear(...) and this is its only gradient: the kernel also emits colsum(dx) = that Linear's bias gradient
            dx, dgamma, dbeta, dxsum = ops.layer_norm_backward(x.v, y.g, gamma.v, eps, want_dx_sum=True)
            x.aux["bias_grad"] = (dxsum, dx)
        else:
            dx, dgamma, dbeta = ops.layer_norm_backward(x.v, y.g, gamma.v, eps)
        if gamma.requires_grad:
            accumulate(gamma, dgamma.reshape(gamma.v.shape))
        if beta.requires_grad:
            accumulate(beta, dbeta.reshape(beta.v.shape))
        if x.requires_grad:
            accumulate(x, dx)
    tape.add(bw)
    return y


# B200_FUSED_ATTENTION: "flash" (default) = forward keeps per-row statistics, backward recomputes the weights — no
# [B,H,S,S] tensor in HBM; "weights" = the round-1 fused kernels that save the weights (what burn-nn's MHA does when
# its `weights` output is asked for); "0" = the scores-GEMM -> softmax -> context-GEMM chain.
_ATTN_MODE = {"1": "flash", "flash": "flash", "weights": "weights", "0": "chain"}[os.environ.get("B200_FUSED_ATTENTION", "flash")]
FUSED_ATTENTION = _ATTN_MODE != "chain"


def attention(tape: Tape, q: Var, k: Var, v: Var, n_heads: int, mask: DeviceTensor | None, causal: bool = False) -> Var:
    """softmax(q·kᵀ/√dk [mask_fill -1e9]) · v on [B,S,d] projections viewed as [B,H,S,dk] (mha.rs:212-311).
    Heads are strided views and the context lands straight in the [B,S,H,dk] layout (no swap_dims copy).
    `causal` says that `mask` is the autoregressive mask (generate_autoregressive_mask): the fused kernels
    then build it from indices and skip the fully masked key blocks — identical results, since a
    -1e9 score contributes exp(-1e9 - max) = 0 exactly.
    Default: flash-style kernels (head dim 64, tf32) — forward saves only per-row softmax statistics, the
    backward recomputes the weights tile by tile and produces dq, dk, dv in two kernels."""
    B, S, d = q.v.shape
    dk = d // n_heads
    def heads(t):
        return t.reshape((B, S, n_heads, dk)).swap_dims(1, 2)
    qh, kh, vh = heads(q.v), heads(k.v), heads(v.v)
    ctx_buf = DeviceTensor.empty((B, S, n_heads, dk))
    ok = dk == 64 and tape.precision == abi.MM_TF32 and S % 4 == 0
    mode = _ATTN_MODE if ok else "chain"
    is_causal = causal and mask is not None
    kmask = None if causal else mask
    scale = 1.0 / math.sqrt(dk)
    w = stats = None
    if mode == "flash":
        _, stats = ops.attention_flash(qh, kh, vh, kmask, scale, -1.0e9, is_causal, out=ctx_buf.swap_dims(1, 2))
    elif mode == "weights":
        _, w = ops.attention(qh, kh, vh, kmask, scale, -1.0e9, is_causal, out=ctx_buf.swap_dims(1, 2), want_weights=True)
    else:
        # scores = mask_fill(q·kᵀ/√dk, mask, -1e9): scaling and mask fill are the GEMM's fused epilogue
        epi = TapeBuilder().op("DIV_F", ("in", 0), ("f", math.sqrt(dk)), out=0 if mask is None else None)
        if mask is not None:
            epi.op("SELECT", "acc", ("f", -1.0e9), ("in", 1), out=0)
        scores = ops.float_matmul(qh, kh.swap_dims(2, 3), tape.precision, epi.build(), () if mask is None else (mask,))
        w = ops.softmax_rows(scores)
        _mm(w, vh, tape.precision, out=ctx_buf.swap_dims(1, 2))
    y = Var(ctx_buf.reshape((B, S, d)), True)
    ctx_h = ctx_buf.swap_dims(1, 2)

    def bw():
        if y.g is None:
            return
        gh = y.g.reshape((B, S, n_heads, dk)).swap_dims(1, 2)          # [B,H,S,dk] view
        dq_buf, dk_buf, dv_buf = (DeviceTensor.empty((B, S, n_heads, dk)) for _ in range(3))
        if mode == "flash":
            ops.attention_flash_backward(gh, qh, kh, vh, ctx_h, stats, kmask, scale, -1.0e9, is_causal,
                                         dq_buf.swap_dims(1, 2), dk_buf.swap_dims(1, 2), dv_buf.swap_dims(1, 2))
        else:
            _mm(w.swap_dims(2, 3), gh, tape.precision, out=dv_buf.swap_dims(1, 2))     # dV = Pᵀ·g
            if mode == "weights":
                # one kernel: dP = g·Vᵀ in TMEM → dS = P∘(dP − rowsum(g∘ctx))/√dk (masked positions have P = 0) → dQ = dS·K
                _, ds = ops.attention_backward(gh, kh, vh, ctx_h, w, scale, is_causal, dq=dq_buf.swap_dims(1, 2))
            else:
                dp = ops.float_matmul(gh, vh.swap_dims(2, 3), tape.precision)
                # softmax backward dS = (dP - sum(dP∘P, -1)) ∘ P, the 1/√dk of the scores and the mask_fill
                # backward (0 where masked) in one row-resident kernel
                ds = ops.softmax_backward(w, dp, mask, math.sqrt(dk))
                _mm(ds, kh, tape.precision, out=dq_buf.swap_dims(1, 2))
            _mm(ds.swap_dims(2, 3), qh, tape.precision, out=dk_buf.swap_dims(1, 2))
        accumulate(q, dq_buf.reshape((B, S, d)))
        accumulate(k, dk_buf.reshape((B, S, d)))
        accumulate(v, dv_buf.reshape((B, S, d)))
    tape.add(bw)
    return y


def embedding(tape: Tape, weight: Var, ids: DeviceTensor) -> Var:
    """select(weight, 0, ids) / embedding_backward = zeros.select_add (ops/modules/base.rs:140-180)."""
    flat = ids.reshape((ids.numel,))
    y = Var(ops.float_select(weight.v, 0, flat).reshape(tuple(ids.shape) + (weight.v.shape[1],)), True)

    def bw():
        if y.g is None or not weight.requires_grad:
            return
        slot = getattr(weight, "grad_slot", None)
        z = slot if slot is not None and weight.g is None else DeviceTensor.empty(weight.v.shape)
        abi.check(abi.load().b200_memset(z.data_ptr(), 0, z.numel * 4, None))
        a, b, c = z.desc(), flat.desc(), y.g.reshape((flat.numel, weight.v.shape[1])).desc()
        abi.check(abi.load().b200_launch_select_add(0, C.byref(a), C.byref(b), C.byref(c), None))
        accumulate(weight, z)
    tape.add(bw)
    return y


FUSED_XENT = os.environ.get("B200_FUSED_XENT", "1") != "0"


def cross_entropy(tape: Tape, logits: Var, targets: DeviceTensor) -> Var:
    """mean(-log_softmax(logits)[target])  (cross_entropy.rs:171-197); logits [N, V], targets i32 [N].
    Fused: one row-resident kernel yields log_softmax[target] and the logits gradient (softmax − onehot)/N
    (written over the logits, which nothing else reads afterwards); unfused: log_softmax → gather → mean,
    and an exp / one-hot tape in backward."""
    n, vsz = logits.v.shape
    if FUSED_XENT and vsz * 4 + 1024 <= 227 * 1024:
        picked, g = ops.softmax_cross_entropy(logits.v, targets.reshape((n,)), 1.0 / n, inplace=True)
        y = Var(ops.float_mul_scalar(ops.float_mean(picked), -1.0), True)

        def bw_fused():
            accumulate(logits, g)
        tape.add(bw_fused)
        return y
    logp = ops.softmax_rows(logits.v, log=True)
    picked = ops.float_gather(1, logp, targets.reshape((n, 1)))
    loss = ops.float_mul_scalar(ops.float_mean(picked), -1.0)
    y = Var(loss, True)

    def bw():
        # d logits = (softmax - onehot) / N   (upstream gradient of the scalar loss is 1)
        cols = DeviceTensor.empty((1, vsz), abi.I32)
        d = cols.desc()
        abi.check(abi.load().b200_launch_arange(C.byref(d), 0, 1, None))
        tb = TapeBuilder()
        tb.op("EQ_I", ("in", 1), ("in", 2))
        tb.op("B2F", "acc", tmp=0)
        tb.op("EXP_F", ("in", 0))
        tb.op("SUB_F", "acc", ("tmp", 0))
        tb.op("MUL_F", "acc", ("f", 1.0 / n), out=0)
        g = _run(tb, [logp, cols.expand((n, vsz)), targets.reshape((n, 1)).expand((n, vsz))], (n, vsz))
        accumulate(logits, g)
    tape.add(bw)
    return y


def mean_square(tape: Tape, x: Var) -> Var:
    sq = ops.float_mul(x.v, x.v)
    y = Var(ops.float_mean(sq), True)

    def bw():
        accumulate(x, ops.float_mul_scalar(x.v, 2.0 / x.v.numel))
    tape.add(bw)
    return y


# ------------------------------------------------------------------ modules (burn-nn)
class Param(Var):
    __slots__ = ("m", "s", "on_grad", "grad_slot", "uses", "_arrived")

    def __init__(self, a: np.ndarray, name: str):
        super().__init__(DeviceTensor.from_numpy(np.ascontiguousarray(a, dtype=np.float32)), True, name)
        self.m = self.s = None  # Adam moments
        self.on_grad = None     # called on every gradient contribution (see ParamArena.ready)
        self.grad_slot = None   # where the gradient should be written (a view of a flat bucket), if any
        self.uses = 1           # how many times the forward pass consumes this parameter (tied weights: 2)
        self._arrived = 0       # contributions seen in the current step


def _uniform(rng, shape, fan_in):
    k = 1.0 / math.sqrt(fan_in)     # Initializer::KaimingUniform{gain 1/√3, fan_out_only false} bound
    return rng.uniform(-k, k, shape).astype(np.float32)


class EncoderLayer:
    def __init__(self, rng, d_model, d_ff, n_heads, idx):
        self.h = n_heads
        p = lambda a, n: Param(a, f"layer{idx}.{n}")
        self.wq, self.bq = p(_uniform(rng, (d_model, d_model), d_model), "wq"), p(_uniform(rng, (d_model,), d_model), "bq")
        self.wk, self.bk = p(_uniform(rng, (d_model, d_model), d_model), "wk"), p(_uniform(rng, (d_model,), d_model), "bk")
        self.wv, self.bv = p(_uniform(rng, (d_model, d_model), d_model), "wv"), p(_uniform(rng, (d_model,), d_model), "bv")
        self.wo, self.bo = p(_uniform(rng, (d_model, d_model), d_model), "wo"), p(_uniform(rng, (d_model,), d_model), "bo")
        self.w1, self.b1 = p(_uniform(rng, (d_model, d_ff), d_model), "w1"), p(_uniform(rng, (d_ff,), d_model), "b1")
        self.w2, self.b2 = p(_uniform(rng, (d_ff, d_model), d_ff), "w2"), p(_uniform(rng, (d_model,), d_ff), "b2")
        self.g1, self.be1 = p(np.ones(d_model), "ln1.gamma"), p(np.zeros(d_model), "ln1.beta")
        self.g2, self.be2 = p(np.ones(d_model), "ln2.gamma"), p(np.zeros(d_model), "ln2.beta")

    def params(self):
        return [self.wq, self.bq, self.wk, self.bk, self.wv, self.bv, self.wo, self.bo, self.w1, self.b1,
                self.w2, self.b2, self.g1, self.be1, self.g2, self.be2]

    def forward(self, tape: Tape, x: Var, mask, causal: bool = False) -> Var:
        q, k, v = linear(tape, x, self.wq, self.bq), linear(tape, x, self.wk, self.bk), linear(tape, x, self.wv, self.bv)
        ctx = attention(tape, q, k, v, self.h, mask, causal)
        x = linear(tape, ctx, self.wo, self.bo, residual=x)       # x + attention output
        x = layer_norm(tape, x, self.g1, self.be1)                 # post-norm (norm_first = false)
        hdn = gelu(tape, linear(tape, x, self.w1, self.b1))
        x = linear(tape, hdn, self.w2, self.b2, residual=x)        # x + feed-forward output
        return layer_norm(tape, x, self.g2, self.be2)


class Encoder:
    """TransformerEncoderConfig::new(d_model, d_ff, n_heads, n_layers), dropout 0."""

    def __init__(self, seed, d_model, d_ff, n_heads, n_layers):
        rng = np.random.default_rng(seed)
        self.layers = [EncoderLayer(rng, d_model, d_ff, n_heads, i) for i in range(n_layers)]

    def params(self):
        return [p for l in self.layers for p in l.params()]

    def forward(self, tape, x, mask=None, causal=False):
        for l in self.layers:
            x = l.forward(tape, x, mask, causal)
        return x


class LanguageModel:
    """examples/text-generation model: token + positional embedding → encoder (causal mask) → vocab head
    (examples/text-generation/src/model.rs:53-100)."""

    def __init__(self, seed, vocab, max_seq, d_model, d_ff, n_heads, n_layers):
        rng = np.random.default_rng(seed)
        self.tok = Param(rng.standard_normal((vocab, d_model)).astype(np.float32), "embedding_token")
        self.pos = Param(rng.standard_normal((max_seq, d_model)).astype(np.float32), "embedding_pos")
        self.enc = Encoder(seed + 1, d_model, d_ff, n_heads, n_layers)
        self.wout = Param(_uniform(rng, (d_model, vocab), d_model), "output.w")
        self.bout = Param(_uniform(rng, (vocab,), d_model), "output.b")

    def params(self):
        return [self.tok, self.pos] + self.enc.params() + [self.wout, self.bout]

    def loss(self, tape, tokens: DeviceTensor, targets: DeviceTensor, pos_ids: DeviceTensor, causal: DeviceTensor):
        B, S = tokens.shape
        x = add(tape, embedding(tape, self.pos, pos_ids), embedding(tape, self.tok, tokens))
        x = scale(tape, x, 0.5)                                   # (pos + tok) / 2, model.rs:66
        h = self.enc.forward(tape, x, causal, causal=True)     # the mask IS generate_autoregressive_mask
        logits = linear(tape, h, self.wout, self.bout)
        return cross_entropy(tape, _reshape(tape, logits, (B * S, logits.v.shape[-1])), targets.reshape((B * S,)))


class FcHead:
    """The fully connected head of examples/mnist (model.rs:26-73 after the conv blocks; BASELINE.json configs[0]):
    fc1(1600,128) → gelu → fc2(128,128) → gelu → fc3(128,10) → cross-entropy.  Dropout is the identity (evaluation
    semantics: a seeded dropout mask is backend-specific).  Weights are [d_in, d_out], KaimingUniform(1/√3)."""
    DIMS = (1600, 128, 128, 10)

    def __init__(self, seed: int):
        rng = np.random.default_rng(seed)
        self.layers = []
        for i, (d_in, d_out) in enumerate(zip(self.DIMS[:-1], self.DIMS[1:])):
            w = Param(_uniform(rng, (d_in, d_out), d_in), f"fc{i + 1}.weight")
            b = Param(_uniform(rng, (d_out,), d_in), f"fc{i + 1}.bias")
            self.layers.append((w, b))

    def params(self):
        return [p for wb in self.layers for p in wb]

    def loss(self, tape: Tape, x: DeviceTensor, targets: DeviceTensor) -> Var:
        h = Var(x, False)
        for i, (w, b) in enumerate(self.layers):
            h = linear(tape, h, w, b)
            if i < len(self.layers) - 1:
                h = gelu(tape, h)
        return cross_entropy(tape, h, targets)


def scale(tape: Tape, x: Var, c: float) -> Var:
    y = Var(ops.float_mul_scalar(x.v, c), True)

    def bw():
        if y.g is not None:
            accumulate(x, ops.float_mul_scalar(y.g, c))
    tape.add(bw)
    return y


def _reshape(tape: Tape, x: Var, shape) -> Var:
    y = Var(x.v.reshape(shape), True)

    def bw():
        if y.g is not None:
            accumulate(x, y.g.reshape(x.v.shape))
    tape.add(bw)
    return y


# ------------------------------------------------------------------ Adam (burn-optim adam.rs:149-210)
class Adam:
    """AdamConfig::new() defaults except lr (crates/burn-optim/src/optim/adam.rs:31-50): β1 0.9, β2 0.999, ε 1e-5.
    The update is the op sequence of AdaptiveMomentum::transform + Adam::step (adam.rs:149-210, :80-84):
        m' = m·β1 + g·(1-β1);  v' = v·β2 + g²·(1-β2)
        u  = (m'·cf) / (√v' + ε_t),  cf = √(1-β2^t)/(1-β1^t),  ε_t = ε·√(1-β2^t);   p' = p - u·lr"""

    def __init__(self, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-5):
        self.lr, self.b1, self.b2, self.eps, self.t = lr, beta1, beta2, eps, 0
        self.coef = None        # device [2]: cf, ε_t — inputs, not immediates, so a captured graph replays

    def coefficients(self, t: int):
        f = np.float32
        bc2s = np.sqrt(f(1.0) - f(self.b2) ** t, dtype=np.float32)
        return f(bc2s / (f(1.0) - f(self.b1) ** t)), f(f(self.eps) * bc2s)

    def advance(self) -> None:
        """Host side of a step: bump t and upload the two time-dependent scalars (8 bytes)."""
        self.t += 1
        if self.coef is None:
            self.coef = DeviceTensor.empty((2,))
        c = np.array(self.coefficients(self.t), dtype=np.float32)
        abi.check(abi.load().b200_memcpy_h2d(self.coef.data_ptr(), c.ctypes.data, 8, None))  # pageable: staged synchronously

    def step(self, params: Sequence[Param]) -> None:
        self.advance()
        self.apply(params)

    def apply(self, params: Sequence[Param]) -> None:
        """One fused elementwise-tape launch per parameter (the op stream burn-fusion would fuse)."""
        for p in params:
            if p.g is None:
                continue
            if p.m is None:
                p.m, p.s = DeviceTensor.empty(p.v.shape), DeviceTensor.empty(p.v.shape)
                lib = abi.load()
                abi.check(lib.b200_memset(p.m.data_ptr(), 0, p.m.numel * 4, None))
                abi.check(lib.b200_memset(p.s.data_ptr(), 0, p.s.numel * 4, None))
            self.update_tape(p.v, p.m, p.s, p.g)

    def apply_arena(self, arena: "ParamArena") -> None:
        """Multi-tensor Adam: one b200_launch_adam per flat bucket instead of one launch per parameter."""
        lib = abi.load()
        if arena.fused:              # the update already ran inside the gradient-sync kernels
            arena.join()
            return
        c = self.coef.desc()
        for b in arena.buckets:      # backward order: early buckets update while late all-reduces still run
            arena.fence(b)
            p, m, s, g = b["p"].desc(), b["m"].desc(), b["s"].desc(), b["g"].desc()
            abi.check(lib.b200_launch_adam(C.byref(p), C.byref(m), C.byref(s), C.byref(g), C.byref(c),
                                           float(self.lr), float(self.b1), float(self.b2), None))
        arena.join()

    def update_tape(self, pv: DeviceTensor, pm: DeviceTensor, ps: DeviceTensor, pg: DeviceTensor) -> None:
        f = np.float32
        omb1, omb2 = float(f(1.0) - f(self.b1)), float(f(1.0) - f(self.b2))
        tb = TapeBuilder()
        tb.op("MUL_F", ("in", 3), ("f", omb1), tmp=0)                      # g·(1-β1)
        tb.op("MUL_F", ("in", 1), ("f", self.b1))
        tb.op("ADD_F", "acc", ("tmp", 0), tmp=1, out=1)                    # m'
        tb.op("MUL_F", ("in", 3), ("in", 3))
        tb.op("MUL_F", "acc", ("f", omb2), tmp=0)                          # g²·(1-β2)
        tb.op("MUL_F", ("in", 2), ("f", self.b2))
        tb.op("ADD_F", "acc", ("tmp", 0), out=2)                           # v'
        tb.op("SQRT_F", "acc")
        tb.op("ADD_F", "acc", ("in", 5), tmp=0)                            # √v' + ε_t
        tb.op("MUL_F", ("tmp", 1), ("in", 4))                              # m'·cf
        tb.op("DIV_F", "acc", ("tmp", 0))
        tb.op("MUL_F", "acc", ("f", self.lr), tmp=0)                       # ·lr
        tb.op("SUB_F", ("in", 0), ("tmp", 0), out=0)                       # p' = p - delta
        c1 = self.coef.slice([(0, 1)]).reshape((1,) * pv.ndim).expand(pv.shape)
        c2 = self.coef.slice([(1, 2)]).reshape((1,) * pv.ndim).expand(pv.shape)
        dv.launch_elemwise(tb.build(), [pv, pm, ps, pg, c1, c2], [pv, pm, ps], pv.shape)

    @staticmethod
    def zero_grad(params):
        for p in params:
            p.g = None


# ------------------------------------------------------------------ flat parameter arena + DDP gradient sync
class ParamArena:
    """Packs parameters, Adam moments and gradients into persistent flat buckets of about
    `bucket_bytes`, in the order backward finalises gradients.

      * gradients: weight-gradient GEMMs and the embedding scatter write their bucket slot directly,
        other gradients are copied in; `p.g` becomes a view of the bucket.
      * DDP (comm given): one ncclAllReduce(avg) per bucket, fired when its last member arrives and
        overlapped with the rest of backward on the collective stream, fenced before the optimizer
        (SURVEY.md §3.4).  The reference issues one collective per parameter
        (crates/burn-cubecl/src/ops/distributed.rs:17-50: 218 tensors in config 5).
      * optimizer: Adam runs as one fused launch per bucket over the flat p/m/v/g arrays
        (multi-tensor Adam) instead of one launch per parameter.
    Persistent storage keeps every address fixed across CUDA-graph replays."""

    @staticmethod
    def peer_bytes(params: Sequence[Param]) -> int:
        """Data-area bytes a PeerGroup needs for these parameters (p and g buckets)."""
        return 2 * 4 * sum((p.v.numel + 3) // 4 * 4 for p in params) + 4096

    def __init__(self, params: Sequence[Param], comm=None, bucket_bytes: int = 32 << 20, peer=None, fused: bool = False):
        """comm: NCCL communicator (one ncclAllReduce per bucket).  peer: a distributed.PeerGroup — the p and g buckets
        then live in the peer region and gradient sync is the peer-memory kernel: `fused=False` all-reduces the bucket
        (Adam runs later, as with NCCL), `fused=True` runs reduce-scatter → Adam on the owned 1/N → all-gather of the
        parameters in ONE kernel per bucket as soon as the bucket is complete (attach the optimizer first)."""
        self.comm = comm
        self.peer, self.fused, self.opt = peer, bool(fused and peer is not None), None
        self.buckets: list[dict] = []
        self.slot: dict[int, dict] = {}
        self._pending: list[dict] = []
        lib = abi.load()
        members, size = [], 0
        order = list(reversed(list(params)))
        for i, p in enumerate(order):
            members.append(p)
            size += (p.v.numel + 3) // 4 * 4              # 16-byte aligned slots
            if size * 4 >= bucket_bytes or i == len(order) - 1:
                flats = {k: DeviceTensor.empty((size,)) for k in ("m", "s")}
                offs = {}
                for k in ("p", "g"):
                    if peer is not None:
                        flats[k], offs[k] = peer.carve(size)
                    else:
                        flats[k] = DeviceTensor.empty((size,))
                for t in flats.values():
                    abi.check(lib.b200_memset(t.data_ptr(), 0, size * 4, None))
                b = dict(flats, n=len(members), arrived=0, done=None, size=size, offs=offs,
                         flag_slot=peer.slot() if peer is not None else -1)
                if comm is not None or peer is not None:   # per-bucket fence: Adam on this bucket waits for its all-reduce only
                    ev = C.c_void_p()
                    abi.check(lib.b200_event_create(C.byref(ev)))
                    b["done"] = ev
                off = 0
                for q in members:
                    view = {k: flats[k].slice([(off, off + q.v.numel)]).reshape(q.v.shape) for k in flats}
                    abi.check(lib.b200_memcpy_d2d(view["p"].data_ptr(), q.v.data_ptr(), q.v.numel * 4, None))
                    q.v, q.m, q.s, q.grad_slot = view["p"], view["m"], view["s"], view["g"]
                    q.on_grad = self.ready
                    self.slot[id(q)] = b
                    off += (q.v.numel + 3) // 4 * 4
                self.buckets.append(b)
                members, size = [], 0
        dv.sync()

    def ready(self, p: Param) -> None:
        """Called by `accumulate` on EVERY gradient contribution to p.  A parameter consumed `p.uses` times
        in the forward pass (tied embedding / output weights, shared layers) is final only after that many
        contributions — counting calls instead would fire the bucket's all-reduce before every member had
        written its slot.  An unexpected extra contribution is an error, never a silently wrong gradient."""
        b = self.slot[id(p)]
        p._arrived += 1
        if p._arrived > p.uses:
            raise RuntimeError(f"parameter {p.name!r} received {p._arrived} gradient contributions this step but "
                               f"declares uses={p.uses}; set Param.uses for shared parameters")
        if p._arrived < p.uses:
            return                      # more contributions to come (accumulate() keeps summing into p.g)
        p._arrived = 0
        if p.g.data_ptr() != p.grad_slot.data_ptr():
            src = p.g if p.g.is_contiguous() else p.g.contiguous()
            abi.check(abi.load().b200_memcpy_d2d(p.grad_slot.data_ptr(), src.data_ptr(), src.numel * 4, None))
            p.g = p.grad_slot
        b["arrived"] += 1
        if b["arrived"] == b["n"]:
            if _SYNC_DEFER:             # experiment: no overlap — every bucket's sync kernel fires after backward, in wait()
                self._pending.append(b)
                b["arrived"] = 0
                return
            self._fire(b)
            b["arrived"] = 0

    def _fire(self, b: dict) -> None:
        if True:
            if self.fused:
                # reduce-scatter → Adam on this rank's 1/N → all-gather of the new parameters, one kernel, beside backward.
                # Safe to update p now: every backward consumer of these parameters has already been launched (a
                # gradient is final only after all of them), and the kernel is fenced behind them.
                o = self.opt
                if o is None or o.coef is None:
                    raise RuntimeError("fused gradient sync needs attach_optimizer(opt) and opt.advance() before backward")
                self.peer.adam(b["offs"]["g"], b["offs"]["p"], b["m"], b["s"], o.coef, b["size"],
                               float(o.lr), float(o.b1), float(o.b2), b["flag_slot"])
                self.peer.mark(b["done"])
            elif self.peer is not None:
                self.peer.all_reduce(b["offs"]["g"], b["size"], b["flag_slot"], mean=True)
                self.peer.mark(b["done"])
            elif self.comm is not None:
                self.comm.all_reduce(b["g"], mean=True)
                abi.check(abi.load().b200_collective_mark(self.comm.handle, b["done"]))

    def attach_optimizer(self, opt: "Adam") -> None:
        self.opt = opt

    def wait(self) -> None:
        """Every gradient has arrived (their all-reduces may still be in flight: the optimizer fences per
        bucket, see Adam.apply_arena)."""
        for b in self.buckets:
            if b["arrived"]:
                raise RuntimeError("a gradient bucket is incomplete: some parameter received no gradient")
        for b in self._pending:
            self._fire(b)
        self._pending.clear()

    def fence(self, b: dict) -> None:
        """The compute stream waits for bucket b's all-reduce only."""
        if self.comm is not None or self.peer is not None:
            abi.check(abi.load().b200_stream_wait_event(None, b["done"]))

    def join(self) -> None:
        """sync_collective: the compute stream rejoins the collective stream completely."""
        if self.peer is not None:
            self.peer.sync()
        elif self.comm is not None:
            self.comm.sync()
