"""Tall, narrow column sums (bias gradients of the training steps) on f32: us and GB/s."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from burn_b200 import _abi as abi, device as dv, ops
from tests import helpers as H
dv.init(0); lib = abi.load()
def timed(fn, it=50):
    for _ in range(5): fn()
    dv.sync()
    e0, e1 = C.c_void_p(), C.c_void_p()
    abi.check(lib.b200_event_create(C.byref(e0))); abi.check(lib.b200_event_create(C.byref(e1)))
    abi.check(lib.b200_event_record(e0, None))
    for _ in range(it): fn()
    abi.check(lib.b200_event_record(e1, None))
    ms = C.c_float(); abi.check(lib.b200_event_elapsed_ms(e0, e1, C.byref(ms)))
    return ms.value / it
for shape in ((16384, 512), (16384, 2048), (8192, 1024), (8192, 4096)):
    x = H.up(np.random.default_rng(0).standard_normal(shape).astype(np.float32))
    t = timed(lambda: ops.float_sum_dim(x, 0))
    gbs = shape[0] * shape[1] * 4 / (t * 1e-3) / 1e9
    print(f"sum_dim(0) {str(shape):14s}: {t*1e3:7.1f} us  {gbs:6.0f} GB/s  ({gbs/6558.7:.2f} of measured peak)")
