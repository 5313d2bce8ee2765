import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from burn_b200 import _abi as abi, device as dv, ops
from tests import helpers as H
dv.init(0)
def rnd(shape, seed): return np.random.default_rng(seed).uniform(-0.5, 0.5, shape).astype(np.float32)
def run(name, fn):
    print("RUN", name, flush=True)
    try:
        err = fn()
        print("  ->", err, flush=True)
    except Exception as e:
        print("  EXC", repr(e)[:300], flush=True)
def case(m, n, k, prec, ta=False, tb=False):
    a, b = rnd((m, k), 1), rnd((k, n), 2)
    da = H.up(np.ascontiguousarray(a.T)).swap_dims(0, 1) if ta else H.up(a)
    db = H.up(np.ascontiguousarray(b.T)).swap_dims(0, 1) if tb else H.up(b)
    got = ops.float_matmul(da, db, prec).numpy()
    ref = a.astype(np.float64) @ b.astype(np.float64)
    bound = np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64)
    return float(np.max(np.abs(got - ref) / (bound + 1e-30)))
which = sys.argv[1:] or ["all"]
cases = {
  "x3_128": lambda: case(128, 128, 128, abi.MM_F32X3),
  "tf32_NT_128": lambda: case(128, 128, 128, abi.MM_TF32, tb=True),     # both K-major in place
  "tf32_NN_128": lambda: case(128, 128, 128, abi.MM_TF32),              # B MN-major
  "tf32_TN_128": lambda: case(128, 128, 128, abi.MM_TF32, ta=True, tb=True),  # A MN-major, B K-major
  "bf16_128": lambda: case(128, 128, 128, abi.MM_BF16),
  "x3_multi": lambda: case(256, 384, 512, abi.MM_F32X3),
  "x3_odd": lambda: case(130, 70, 33, abi.MM_F32X3),
  "nt_n2": lambda: case(128, 256, 128, abi.MM_TF32, tb=True),
  "nt_m2": lambda: case(256, 128, 128, abi.MM_TF32, tb=True),
  "nt_k512": lambda: case(128, 128, 512, abi.MM_TF32, tb=True),
  "nt_k2048": lambda: case(128, 128, 2048, abi.MM_TF32, tb=True),
  "nt_m2n2": lambda: case(256, 256, 128, abi.MM_TF32, tb=True),
  "x3_big": lambda: case(1024, 1024, 1024, abi.MM_F32X3),
}
for name, fn in cases.items():
    if which == ["all"] or name in which:
        run(name, fn)
