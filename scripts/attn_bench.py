"""Fused attention forward vs the unfused chain on the configs[4] shape [8,16,1024,64]."""
import os, sys, ctypes as C, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from burn_b200 import _abi as abi, device as dv, ops
from burn_b200.device import DeviceTensor, TapeBuilder
from tests import helpers as H
dv.init(0); lib = abi.load()
B, Hh, S, dk = (int(x) for x in os.environ.get("ATTN_SHAPE", "8,16,1024,64").split(","))
rng = np.random.default_rng(0)
heads = lambda: H.up((rng.standard_normal((B, S, Hh * dk)) * 0.5).astype(np.float32)).reshape((B, S, Hh, dk)).swap_dims(1, 2)
q, k, v = heads(), heads(), heads()
mask = H.up(np.triu(np.ones((S, S), dtype=bool), k=1)[None, None])
ctx = DeviceTensor.empty((B, S, Hh, dk))
def timed(fn, it=10):
    for _ in range(2): fn()
    dv.sync()
    e0, e1 = C.c_void_p(), C.c_void_p()
    abi.check(lib.b200_event_create(C.byref(e0))); abi.check(lib.b200_event_create(C.byref(e1)))
    abi.check(lib.b200_event_record(e0, None))
    for _ in range(it): fn()
    abi.check(lib.b200_event_record(e1, None))
    ms = C.c_float(); abi.check(lib.b200_event_elapsed_ms(e0, e1, C.byref(ms)))
    return ms.value / it * 1e3
def chain():
    epi = TapeBuilder().op("DIV_F", ("in", 0), ("f", 8.0)).op("SELECT", "acc", ("f", -1.0e9), ("in", 1), out=0).build()
    sc = ops.float_matmul(q, k.swap_dims(2, 3), abi.MM_TF32, epi, (mask,))
    w = ops.softmax_rows(sc)
    return ops.float_matmul(w, v, abi.MM_TF32)
print(f"unfused chain (scores GEMM+mask epilogue, softmax, context GEMM): {timed(chain):8.1f} us")
for name, kw in (("fused, causal flag, weights", dict(is_causal=True, want_weights=True)), ("fused, causal flag, no weights", dict(is_causal=True)),
                 ("fused, mask tensor, weights", dict(mask=mask, want_weights=True)), ("fused, no mask, no weights", dict())):
    t = timed(lambda: ops.attention(q, k, v, kw.get("mask"), 0.125, -1.0e9, kw.get("is_causal", False), out=ctx.swap_dims(1, 2),
                                    want_weights=kw.get("want_weights", False)))
    print(f"{name:34s}: {t:8.1f} us")
# ---- flash variants (no [B,H,S,S] tensor): forward + backward
g = heads()
dqb, dkb, dvb = (DeviceTensor.empty((B, S, Hh, dk)) for _ in range(3))
for name, causal, m in (("flash fwd, causal", True, None), ("flash fwd, no mask", False, None), ("flash fwd, mask tensor", False, mask)):
    t = timed(lambda: ops.attention_flash(q, k, v, m, 0.125, -1.0e9, causal, out=ctx.swap_dims(1, 2)))
    print(f"{name:34s}: {t:8.1f} us")
for name, causal, m in (("flash bwd (dq + dkv), causal", True, None), ("flash bwd (dq + dkv), no mask", False, None)):
    _, stats = ops.attention_flash(q, k, v, m, 0.125, -1.0e9, causal, out=ctx.swap_dims(1, 2))
    t = timed(lambda: ops.attention_flash_backward(g, q, k, v, ctx.swap_dims(1, 2), stats, m, 0.125, -1.0e9, causal,
                                                   dqb.swap_dims(1, 2), dkb.swap_dims(1, 2), dvb.swap_dims(1, 2)))
    print(f"{name:34s}: {t:8.1f} us")
