import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from burn_b200 import _abi as abi, device as dv
from burn_b200.device import DeviceTensor, TapeBuilder
from tests import helpers as H
from tests.test_elemwise_gpu import bench_chain_tape
dv.init(0)
n = 8192
rng = np.random.default_rng(0)
a = rng.uniform(-1, 1, (n, n)).astype(np.float32)
da, db, dc = H.up(a), H.up(rng.uniform(-1, 1, (n, n)).astype(np.float32)), H.up(rng.uniform(-1, 1, (n, n)).astype(np.float32))
dm = H.up(a < 0)
out = DeviceTensor.empty((n, n))
which = sys.argv[1] if len(sys.argv) > 1 else "all"
for _ in range(3):
    if which in ("all", "mul"):
        dv.launch_elemwise(TapeBuilder().op("MUL_F", ("in", 0), ("f", 2.0), out=0).build(), [da], [out], (n, n))
    if which in ("all", "chain"):
        dv.launch_elemwise(bench_chain_tape().build(), [da, db, dc, dm], [out], (n, n))
dv.sync()
