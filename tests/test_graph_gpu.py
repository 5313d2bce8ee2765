"""GPU: CUDA-graph capture entry points (b200_graph_begin/end/launch): a launch sequence with buffers
allocated and freed inside the capture replays with one call and sees the current contents of the
persistent inputs."""
import numpy as np
import pytest

from burn_b200 import _abi as abi
from burn_b200 import device as dv
from burn_b200 import ops
from burn_b200.device import DeviceTensor
from oracle import oracle
from tests import helpers as H

pytestmark = pytest.mark.gpu


def test_captured_sequence_replays_on_new_inputs(dev):
    rng = np.random.default_rng(0)
    a = rng.uniform(-1, 1, (256, 512)).astype(np.float32)
    w = rng.uniform(-1, 1, (512, 128)).astype(np.float32)
    x, dw = H.up(a), H.up(w)                      # persistent buffers (allocated before the capture)
    out = DeviceTensor.empty((256, 1))

    def body():
        y = ops.gelu(ops.float_matmul(x, dw, abi.MM_F32X3))        # temporaries live and die inside the capture
        s = ops.float_sum_dim(y, 1)
        abi.check(abi.load().b200_memcpy_d2d(out.data_ptr(), s.data_ptr(), 256 * 4, None))

    body()
    dv.sync()
    eager = out.numpy().copy()
    with dv.Graph.capture() as g:
        body()
    assert g.kernel_nodes >= 3
    g.launch()
    dv.sync()
    assert np.array_equal(out.numpy(), eager)
    # new contents, same addresses
    a2 = rng.uniform(-1, 1, (256, 512)).astype(np.float32)
    abi.check(abi.load().b200_memcpy_h2d(x.data_ptr(), a2.ctypes.data, a2.nbytes, None))
    g.launch()
    dv.sync()
    want = oracle.float_sum_dim(oracle.gelu(oracle.float_matmul(a2[None], w[None])[0]), 1)
    H.assert_close(out.numpy(), want, 2e-5, 2e-4, "graph replay on new inputs")
    before = int(abi.load().b200_launch_count())
    g.launch()
    assert int(abi.load().b200_launch_count()) - before == g.kernel_nodes
    dv.sync()
    g.destroy()
