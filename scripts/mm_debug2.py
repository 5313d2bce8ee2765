import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from burn_b200 import _abi as abi, device as dv, ops
from tests import helpers as H
dv.init(0)
n = 128
a = np.eye(n, dtype=np.float32)
kk, nn = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
for name, b in (("b=k", kk.astype(np.float32)), ("b=n", nn.astype(np.float32))):
    got = ops.float_matmul(H.up(a), H.up(b), abi.MM_TF32).numpy()   # A K-major, B MN-major
    print(name, "equal to b:", np.array_equal(got, b), " equal to b.T:", np.array_equal(got, b.T))
    np.set_printoptions(linewidth=250)
    print(got[:10, :40].astype(int))
    print(got[30:36, :40].astype(int))
