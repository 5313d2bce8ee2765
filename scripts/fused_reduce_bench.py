"""sum_dim(gelu(a*b), axis) on [8192, 8192] f32: fused read tape vs HBM (8 B/elem algorithmic)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from burn_b200 import _abi as abi, device as dv
from burn_b200.device import DeviceTensor, TapeBuilder
from tests import helpers as H
dv.init(0); lib = abi.load()
n = 8192
rng = np.random.default_rng(0)
a = H.up(rng.uniform(-1, 1, (n, n)).astype(np.float32)); b = H.up(rng.uniform(-1, 1, (n, n)).astype(np.float32))
tb = TapeBuilder().op("MUL_F", ("in", 0), ("in", 1), tmp=0); H.gelu_tape(tb, ("tmp", 0)); tape = tb.build()
for axis in (1, 0):
    keep = [n, n]; keep[axis] = 1
    out = DeviceTensor.empty(keep)
    fn = lambda: dv.launch_reduce(abi.RED_SUM, axis, (n, n), [a, b], [out], read=tape)
    for _ in range(3): fn()
    dv.sync()
    e0, e1 = C.c_void_p(), C.c_void_p()
    abi.check(lib.b200_event_create(C.byref(e0))); abi.check(lib.b200_event_create(C.byref(e1)))
    abi.check(lib.b200_event_record(e0, None))
    for _ in range(20): fn()
    abi.check(lib.b200_event_record(e1, None))
    ms = C.c_float(); abi.check(lib.b200_event_elapsed_ms(e0, e1, C.byref(ms)))
    t = ms.value / 20
    print(f"sum_dim(gelu(a*b), {axis}) [8192,8192]: {t*1e3:.1f} us  {n*n*8/(t*1e-3)/1e9:.0f} GB/s  ({n*n*8/(t*1e-3)/1e9/6558.7:.2f} of measured peak)")
