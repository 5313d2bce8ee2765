"""Plain fast reductions on [8192, 8192] f32 (4 B/elem in + outputs): sum / mean / max / argmax on both axes."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from burn_b200 import _abi as abi, device as dv, ops
from tests import helpers as H
dv.init(0); lib = abi.load()
n = 8192
x = H.up(np.random.default_rng(0).standard_normal((n, n)).astype(np.float32))
def timed(fn, it=20):
    for _ in range(3): fn()
    dv.sync()
    e0, e1 = C.c_void_p(), C.c_void_p()
    abi.check(lib.b200_event_create(C.byref(e0))); abi.check(lib.b200_event_create(C.byref(e1)))
    abi.check(lib.b200_event_record(e0, None))
    for _ in range(it): fn()
    abi.check(lib.b200_event_record(e1, None))
    ms = C.c_float(); abi.check(lib.b200_event_elapsed_ms(e0, e1, C.byref(ms)))
    return ms.value / it
for _ in range(200): ops.float_sum_dim(x, 1)      # clocks up before the first measurement
dv.sync()
for name, fn in (("sum_dim(1)", lambda: ops.float_sum_dim(x, 1)), ("sum_dim(0)", lambda: ops.float_sum_dim(x, 0)),
                 ("max_dim(1)", lambda: ops.float_max_dim(x, 1)), ("argmax(1)", lambda: ops.float_argmax(x, 1)),
                 ("argmax(0)", lambda: ops.float_argmax(x, 0)), ("argmin(1)", lambda: ops.float_argmin(x, 1))):
    t = timed(fn)
    gbs = (n * n * 4 + n * 4) / (t * 1e-3) / 1e9
    print(f"{name:12s} [8192,8192]: {t*1e3:7.1f} us  {gbs:6.0f} GB/s  ({gbs/6558.7:.2f} of measured peak)")
