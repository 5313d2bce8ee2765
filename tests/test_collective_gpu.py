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


# ---------------------------------------------------------------- peer-memory collectives (burn_b200/csrc/peer.cu)
def _adam_ref(lib, p, m, v, g, coef):
    import ctypes as C
    a, b, c, d, e = p.desc(), m.desc(), v.desc(), g.desc(), coef.desc()
    abi.check(lib.b200_launch_adam(C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(e), 1e-3, 0.9, 0.999, None))


def test_peer_group_of_one_matches_the_plain_kernels_and_replays_in_a_graph(dev):
    """World of 1 on one GPU: the flag / epoch protocol, the Sum / Mean identity and the fused Adam — which must equal
    b200_launch_adam bit for bit — launch after launch and through CUDA-graph replay (epochs advance on the device)."""
    from burn_b200 import device as dv
    from burn_b200.device import DeviceTensor
    from burn_b200.distributed import PeerGroup
    lib = abi.load()
    n = (1 << 20) + 8
    grp = PeerGroup(0, 1, 4 * 2 * n + 4096)
    g, g_off = grp.carve(n)
    p, p_off = grp.carve(n)
    s_ar, s_adam = grp.slot(), grp.slot()
    rng = np.random.default_rng(0)
    x = rng.standard_normal(n).astype(np.float32)
    abi.check(lib.b200_memcpy_h2d(g.data_ptr(), x.ctypes.data, n * 4, None))
    for mean in (False, True, True):
        grp.all_reduce(g_off, n, s_ar, mean=mean)
        grp.sync()
    H.assert_exact(g.numpy(), x)
    p0 = rng.standard_normal(n).astype(np.float32)
    abi.check(lib.b200_memcpy_h2d(p.data_ptr(), p0.ctypes.data, n * 4, None))
    zeros = np.zeros(n, np.float32)
    m, v, rp, rm, rv = (DeviceTensor.from_numpy(a) for a in (zeros, zeros, p0, zeros, zeros))
    coef = DeviceTensor.from_numpy(np.array([0.31622776, 3.1622776e-7], dtype=np.float32))
    for it in range(3):
        grp.adam(g_off, p_off, m, v, coef, n, 1e-3, 0.9, 0.999, s_adam)
        grp.sync()
        _adam_ref(lib, rp, rm, rv, g, coef)
    H.assert_exact(p.numpy(), rp.numpy())
    H.assert_exact(m.numpy(), rm.numpy())
    H.assert_exact(v.numpy(), rv.numpy())
    with dv.Graph.capture() as graph:
        grp.adam(g_off, p_off, m, v, coef, n, 1e-3, 0.9, 0.999, s_adam)
        grp.sync()
    for _ in range(4):
        graph.launch()
        _adam_ref(lib, rp, rm, rv, g, coef)
    dv.sync()
    H.assert_exact(p.numpy(), rp.numpy())
    H.assert_exact(v.numpy(), rv.numpy())
    graph.destroy()
    grp.close()


def _n_gpus():
    import ctypes as C
    n = C.c_int32(0)
    return n.value if abi.load().b200_device_count(C.byref(n)) == 0 and n.value else n.value


needs2 = pytest.mark.skipif("_n_gpus() < 2", reason="needs two GPUs (run with gpurun --gpus 2)")


@needs2
def test_one_thread_drives_all_devices_peer_and_nccl(dev):
    """The reference's DDP shape: ONE process, ONE thread issuing the all_reduce of every device
    (crates/burn-backend/src/backend/distributed/server.rs:100-139).  Sum / Mean on device-dependent data
    (tests/tensor/distributed.rs:26-62) through (a) the peer-memory kernels and (b) ncclGroupStart/End over
    communicators from ncclCommInitAll; then the fused reduce-scatter + Adam + all-gather against the unfused update."""
    import ctypes as C
    lib = abi.load()
    nd = min(_n_gpus(), 4)
    n = 1 << 18
    devices = (C.c_int32 * nd)(*range(nd))
    flag = int(lib.b200_peer_flag_bytes())
    groups = (C.c_void_p * nd)()
    abi.check(lib.b200_peer_group_create_local(groups, devices, nd, flag + 4 * 5 * n + 64))
    base = [int(lib.b200_peer_data(groups[i])) for i in range(nd)]
    # data area layout (elements): g [0,n)  p [n,2n)  m [2n,3n)  v [3n,4n)  coef [4n, 4n+4)
    data = [np.random.default_rng([31, r]).uniform(0, 10, n).astype(np.float32) for r in range(nd)]
    acc = data[0].copy()
    for r in range(1, nd):
        acc = acc + data[r]

    def put(i, off, a):
        abi.check(lib.b200_memcpy_h2d(base[i] + off * 4, a.ctypes.data, a.nbytes, None))

    def get(i, off, count):
        out = np.empty(count, np.float32)
        abi.check(lib.b200_memcpy_d2h(out.ctypes.data, base[i] + off * 4, count * 4, None))
        abi.check(lib.b200_stream_sync(None))
        return out
    for mean in (False, True):
        for i in range(nd):
            put(i, 0, data[i])
        abi.check(lib.b200_stream_sync(None))
        for i in range(nd):                                   # one thread, device after device: launches are asynchronous
            abi.check(lib.b200_launch_peer_all_reduce(groups[i], 0, n, abi.REDUCE_MEAN if mean else abi.REDUCE_SUM, 0, None))
        for i in range(nd):
            abi.check(lib.b200_peer_host_sync(groups[i]))
        want = acc / np.float32(nd) if mean else acc
        for i in range(nd):
            H.assert_exact(get(i, 0, n), want)
    # fused Adam
    p0 = np.random.default_rng(32).standard_normal(n).astype(np.float32)
    coef = np.array([0.31622776, 3.1622776e-7, 0, 0], dtype=np.float32)
    for i in range(nd):
        put(i, 0, data[i]); put(i, n, p0); put(i, 2 * n, np.zeros(2 * n, np.float32)); put(i, 4 * n, coef)
    abi.check(lib.b200_stream_sync(None))
    for i in range(nd):
        abi.check(lib.b200_launch_peer_adam(groups[i], 0, n, base[i] + 2 * n * 4, base[i] + 3 * n * 4, base[i] + 4 * n * 4,
                                            n, 1e-3, 0.9, 0.999, 1, None))
    for i in range(nd):
        abi.check(lib.b200_peer_host_sync(groups[i]))
    from burn_b200.device import DeviceTensor
    rp, rm, rv = DeviceTensor.from_numpy(p0), DeviceTensor.from_numpy(np.zeros(n, np.float32)), DeviceTensor.from_numpy(np.zeros(n, np.float32))
    _adam_ref(lib, rp, rm, rv, DeviceTensor.from_numpy(acc / np.float32(nd)), DeviceTensor.from_numpy(coef[:2]))
    want_p = rp.numpy()
    for i in range(nd):
        H.assert_exact(get(i, n, n), want_p)                  # every device holds the same, correct parameters
    for i in range(nd):
        abi.check(lib.b200_peer_group_destroy(groups[i]))
    # (b) NCCL: one group call for all devices
    comms = (C.c_void_p * nd)()
    abi.check(lib.b200_comm_init_all(comms, devices, nd))
    groups2 = (C.c_void_p * nd)()
    abi.check(lib.b200_peer_group_create_local(groups2, devices, nd, flag + 4 * n + 64))   # just device-local buffers
    base2 = [int(lib.b200_peer_data(groups2[i])) for i in range(nd)]
    for mean in (False, True):
        for i in range(nd):
            abi.check(lib.b200_memcpy_h2d(base2[i], data[i].ctypes.data, n * 4, None))
        abi.check(lib.b200_stream_sync(None))
        ptrs = (C.c_void_p * nd)(*base2)
        abi.check(lib.b200_all_reduce_group(comms, ptrs, n, nd, abi.F32, abi.REDUCE_MEAN if mean else abi.REDUCE_SUM, None))
        for i in range(nd):
            abi.check(lib.b200_comm_host_sync(comms[i]))
        ref = np.sum(np.stack(data).astype(np.float64), axis=0) / (nd if mean else 1)
        for i in range(nd):
            out = np.empty(n, np.float32)
            abi.check(lib.b200_memcpy_d2h(out.ctypes.data, base2[i], n * 4, None))
            abi.check(lib.b200_stream_sync(None))
            assert np.all(np.abs(out - ref) <= 1e-6 * np.abs(ref) * nd + 1e-6)
    for i in range(nd):
        abi.check(lib.b200_comm_destroy(comms[i]))
        abi.check(lib.b200_peer_group_destroy(groups2[i]))
