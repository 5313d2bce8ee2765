"""GPU: b200_all_reduce (NCCL inside libburn_b200.so).  Single-rank communicator on one GPU —
Sum and Mean over a world of 1 must be the identity, in place, ordered after the producer stream
(the multi-rank run is exercised by `bench.py --gpus N` and scripts/allreduce_check.py).
Mirrors crates/burn-backend-tests/tests/tensor/distributed.rs:10-62."""
import numpy as np
import pytest

from burn_b200 import _abi as abi
from tests import helpers as H

pytestmark = pytest.mark.gpu


def test_single_rank_all_reduce_is_identity_and_ordered(dev):
    from burn_b200 import ops
    from burn_b200.distributed import Communicator
    comm = Communicator(0, 1)
    x = np.random.default_rng(0).standard_normal((1 << 20,)).astype(np.float32)
    t = H.up(x)
    for _ in range(5):
        t = ops.float_mul_scalar(t, 2.0)          # producer work on the default stream
        comm.all_reduce(t, mean=True)              # must wait for it (event fence)
        comm.sync()                                # consumer waits for the collective
    H.assert_exact(t.numpy(), x * np.float32(32.0))
    parts = [H.up(np.full((1000 + i,), float(i), dtype=np.float32)) for i in range(8)]
    comm.all_reduce_bucket(parts, mean=False)
    comm.sync()
    for i, p in enumerate(parts):
        H.assert_exact(p.numpy(), np.full((1000 + i,), float(i), dtype=np.float32))
    comm.close()
