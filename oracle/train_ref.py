"""CPU restatement (numpy, f32 storage) of a training step of the MNIST example's fully connected head —
BASELINE.json configs[0], the reference's own CPU-runnable case.  TEST INFRASTRUCTURE: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.  Parity unpinned by reference fixtures (the
reference holds no golden loss values for this model); it is checked op by op instead: every primitive below is the
pinned oracle's (oracle.py) or a NumPy/OpenBLAS sgemm standing in for matrixmultiply::sgemm
(crates/burn-ndarray/src/ops/matmul.rs:35-61), composed in the order burn-nn / burn-autodiff / burn-optim record them.

Model (examples/mnist/src/model.rs:26-73, the part after the conv blocks; dropout is the identity here — seeded
dropout masks are not reproducible across backends, stated in DESIGN.md):
    x[B,1600] -> fc1(1600,128) -> gelu -> fc2(128,128) -> gelu -> fc3(128,10) -> cross-entropy(mean)
Backward: burn-autodiff's rules — linear: dX = dY·Wᵀ, dW = Xᵀ·dY, db = sum_dim(dY, 0) (crates/burn-autodiff/src/ops/
tensor.rs matmul/add backward); gelu: B::gelu_backward, the tanh-approximation derivative the trait default uses
(crates/burn-backend/src/backend/ops/activation.rs:98-128); cross-entropy: (softmax − onehot)/N
(crates/burn-nn/src/loss/cross_entropy.rs:171-197 under autodiff).  Optimizer: Adam, burn-optim defaults
(crates/burn-optim/src/optim/adam.rs:31-50,80-84,149-210), f32 state.
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32
DIMS = (1600, 128, 128, 10)


def init_params(seed: int):
    """LinearConfig default initializer: KaimingUniform(gain 1/sqrt3) = U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for the
    weight and the bias (crates/burn-nn/src/modules/linear.rs:21,60-75).  Same stream as burn_b200.train.FcHead."""
    rng = np.random.default_rng(seed)
    ps = []
    for d_in, d_out in zip(DIMS[:-1], DIMS[1:]):
        k = 1.0 / math.sqrt(d_in)
        ps.append(rng.uniform(-k, k, (d_in, d_out)).astype(F))
        ps.append(rng.uniform(-k, k, (d_out,)).astype(F))
    return ps


def gelu(x):
    from oracle import oracle
    return oracle.gelu(x)


def gelu_backward(x, g):
    """The 20-op default of ActivationOps::gelu_backward, f32 op by op (tanh via f64 like the oracle)."""
    from oracle import oracle as o
    x3 = o.float_powf_scalar(x, 3.0)
    c1 = o.float_mul_scalar(x3, 0.0356774)
    c2 = o.float_mul_scalar(x, 0.797885)
    c3 = o.float_mul_scalar(x3, 0.0535161)
    c4 = o.float_mul_scalar(x, 0.398942)
    tanh = o.float_tanh(o.float_add(c1, c2))
    inner2 = o.float_add(c3, c4)
    sech = o.float_add_scalar(o.float_neg(o.float_mul(tanh, tanh)), 1.0)
    y1 = o.float_mul_scalar(tanh, 0.5)
    y2 = o.float_add_scalar(o.float_mul(inner2, sech), 0.5)
    return o.float_mul(o.float_add(y1, y2), g)


def forward_backward(params, x, targets):
    """Returns (loss f32, grads list) for one batch; x [B,1600] f32, targets [B] int."""
    w1, b1, w2, b2, w3, b3 = params
    n = x.shape[0]
    z1 = (x @ w1 + b1).astype(F)
    h1 = gelu(z1)
    z2 = (h1 @ w2 + b2).astype(F)
    h2 = gelu(z2)
    logits = (h2 @ w3 + b3).astype(F)
    m = logits.max(axis=1, keepdims=True)
    sh = (logits - m).astype(F)
    lse = np.log(np.exp(sh).astype(F).sum(axis=1, keepdims=True, dtype=F)).astype(F)
    logp = (sh - lse).astype(F)
    loss = F(-logp[np.arange(n), targets].mean(dtype=F))
    dlog = np.exp(logp).astype(F)
    dlog[np.arange(n), targets] -= F(1.0)
    dlog = (dlog * F(1.0 / n)).astype(F)
    gw3, gb3 = (h2.T @ dlog).astype(F), dlog.sum(axis=0, dtype=F)
    dh2 = (dlog @ w3.T).astype(F)
    dz2 = gelu_backward(z2, dh2)
    gw2, gb2 = (h1.T @ dz2).astype(F), dz2.sum(axis=0, dtype=F)
    dh1 = (dz2 @ w2.T).astype(F)
    dz1 = gelu_backward(z1, dh1)
    gw1, gb1 = (x.T @ dz1).astype(F), dz1.sum(axis=0, dtype=F)
    return loss, [gw1, gb1, gw2, gb2, gw3, gb3]


class Adam:
    def __init__(self, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-5):
        self.lr, self.b1, self.b2, self.eps, self.t = F(lr), F(beta1), F(beta2), F(eps), 0
        self.m = self.s = None

    def step(self, params, grads):
        if self.m is None:
            self.m = [np.zeros_like(p) for p in params]
            self.s = [np.zeros_like(p) for p in params]
        self.t += 1
        bc2s = np.sqrt(F(1.0) - self.b2 ** self.t, dtype=F)
        cf, eps_t = F(bc2s / (F(1.0) - self.b1 ** self.t)), F(self.eps * bc2s)
        for i, (p, g) in enumerate(zip(params, grads)):
            self.m[i] = (self.m[i] * self.b1 + g * (F(1.0) - self.b1)).astype(F)
            self.s[i] = (self.s[i] * self.b2 + (g * g) * (F(1.0) - self.b2)).astype(F)
            u = ((self.m[i] * cf) / (np.sqrt(self.s[i]) + eps_t)).astype(F)
            params[i] = (p - u * self.lr).astype(F)


_TEACHER = {}


def batch(seed: int, step: int, n: int = 64):
    """Synthetic stand-in for a post-conv MNIST batch (no dataset offline): relu-sparse features, a fresh batch every
    step, labels from a fixed random linear teacher so that the loss curve actually descends."""
    if seed not in _TEACHER:
        _TEACHER[seed] = np.random.default_rng([seed, 1 << 20]).standard_normal((DIMS[0], DIMS[-1]))
    rng = np.random.default_rng([seed, step])
    x = np.maximum(rng.standard_normal((n, DIMS[0])), 0.0).astype(F)
    return x, np.argmax(x.astype(np.float64) @ _TEACHER[seed], axis=1).astype(np.int64)


def train(seed: int, steps: int, lr: float = 1e-3, n: int = 64):
    params, opt, losses = init_params(seed), Adam(lr), []
    for s in range(steps):
        x, t = batch(seed, s, n)
        loss, grads = forward_backward(params, x, t)
        opt.step(params, grads)
        losses.append(float(loss))
    return losses, params
