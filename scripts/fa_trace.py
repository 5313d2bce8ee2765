"""One flash backward on [8,16,1024,64] (no mask) with the -DFA_TRACE library: block 0 prints its pipeline timeline."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from burn_b200 import _abi as abi, device as dv, ops
from burn_b200.device import DeviceTensor
from tests import helpers as H
dv.init(0); lib = abi.load()
B, Hh, S, dk = 8, 16, 1024, 64
rng = np.random.default_rng(0)
heads = lambda: H.up((rng.standard_normal((B, S, Hh * dk)) * 0.5).astype(np.float32)).reshape((B, S, Hh, dk)).swap_dims(1, 2)
q, k, v, g = heads(), heads(), heads(), heads()
ctx = DeviceTensor.empty((B, S, Hh, dk))
dqb, dkb, dvb = (DeviceTensor.empty((B, S, Hh, dk)) for _ in range(3))
causal = len(sys.argv) > 1 and sys.argv[1] == "causal"
_, stats = ops.attention_flash(q, k, v, None, 0.125, -1.0e9, causal, out=ctx.swap_dims(1, 2))
for _ in range(2):
    ops.attention_flash_backward(g, q, k, v, ctx.swap_dims(1, 2), stats, None, 0.125, -1.0e9, causal,
                                 dqb.swap_dims(1, 2), dkb.swap_dims(1, 2), dvb.swap_dims(1, 2))
    dv.sync()
