"""softmax cross-entropy on the configs[4] head [8192, 50260]: fused kernel vs 8 B/elem at HBM speed."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from burn_b200 import _abi as abi, device as dv, ops
from burn_b200.device import DeviceTensor
from tests import helpers as H
dv.init(0); lib = abi.load()
n, v = 8192, 50260
rng = np.random.default_rng(0)
x = H.up(rng.standard_normal((n, v)).astype(np.float32)); t = H.up(rng.integers(0, v, n).astype(np.int32))
fn = lambda: ops.softmax_cross_entropy(x, t, 1.0 / n, inplace=False)
for _ in range(2): fn()
dv.sync()
e0, e1 = C.c_void_p(), C.c_void_p()
abi.check(lib.b200_event_create(C.byref(e0))); abi.check(lib.b200_event_create(C.byref(e1)))
abi.check(lib.b200_event_record(e0, None))
for _ in range(10): fn()
abi.check(lib.b200_event_record(e1, None))
ms = C.c_float(); abi.check(lib.b200_event_elapsed_ms(e0, e1, C.byref(ms)))
tt = ms.value / 10
print(f"softmax_cross_entropy [{n},{v}]: {tt*1e3:.0f} us  {n*v*8/(tt*1e-3)/1e9:.0f} GB/s ({n*v*8/(tt*1e-3)/1e9/6558.7:.2f} of peak)")
