"""ncu targets: one 2048^3 and one 16384^3 GEMM per precision (bf16 first)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from burn_b200 import _abi as abi, device as dv
from burn_b200.device import DeviceTensor
from tests import helpers as H
dv.init(0); lib = abi.load()
n = int(sys.argv[1])
rng = np.random.default_rng(1)
a = rng.uniform(-0.5, 0.5, (n, n)).astype(np.float32); b = rng.uniform(-0.5, 0.5, (n, n)).astype(np.float32)
for prec, bf in ((abi.MM_BF16, True), (abi.MM_TF32, False)):
    da = DeviceTensor.from_bf16_of(a) if bf else H.up(a)
    db = (DeviceTensor.from_bf16_of(b) if bf else H.up(b)).swap_dims(0, 1)
    out = DeviceTensor.empty((n, n))
    ad, bd, cd = da.desc(), db.desc(), out.desc()
    for _ in range(2):
        abi.check(lib.b200_launch_matmul(C.byref(ad), C.byref(bd), C.byref(cd), prec, None, None, 0, None, 0, None))
    dv.sync()
