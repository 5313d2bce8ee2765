"""Under torchrun: value checks of the peer-memory collectives (bench.peer_value_check), then time per 256 MiB f32 bucket
for ncclAllReduce(avg), the peer-memory all-reduce kernel, NCCL + multi-tensor Adam, and the fused
reduce-scatter + Adam(1/N) + all-gather kernel — alone on the GPU (no competing compute)."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from burn_b200 import _abi as abi, device as dv
from burn_b200.device import DeviceTensor
from burn_b200.distributed import Communicator, PeerGroup
import bench
dv.init(local)
lib = abi.load()
bench.peer_value_check(rank, world, local)
if rank == 0:
    print("peer value check ok", flush=True)
n = 64 << 20
comm = Communicator(rank, world, device=torch.device("cuda", local))
grp = PeerGroup(rank, world, 2 * 4 * n + 4096, device=torch.device("cuda", local))
g, g_off = grp.carve(n)
p, p_off = grp.carve(n)
m, v = DeviceTensor.empty((n,)), DeviceTensor.empty((n,))
for t in (m, v):
    abi.check(lib.b200_memset(t.data_ptr(), 0, n * 4, None))
coef = DeviceTensor.from_numpy(np.array([0.31622776, 3.1622776e-7], dtype=np.float32))
s_ar, s_adam = grp.slot(), grp.slot()
big = DeviceTensor.empty((n,))
abi.check(lib.b200_memset(big.data_ptr(), 0, n * 4, None))


def adam_local():
    a, b, c, d, e = p.desc(), m.desc(), v.desc(), big.desc(), coef.desc()
    abi.check(lib.b200_launch_adam(C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(e), 1e-3, 0.9, 0.999, None))


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    dv.sync(); dist.barrier()
    e0, e1 = C.c_void_p(), C.c_void_p()
    lib.b200_event_create(C.byref(e0)); lib.b200_event_create(C.byref(e1))
    lib.b200_event_record(e0, None)
    for _ in range(iters):
        fn()
    lib.b200_event_record(e1, None)
    ms = C.c_float(); abi.check(lib.b200_event_elapsed_ms(e0, e1, C.byref(ms)))
    t = torch.tensor([ms.value], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / iters


def nccl_ar():
    comm.all_reduce(big, mean=True); comm.sync()


def peer_ar():
    grp.all_reduce(g_off, n, s_ar, mean=True); grp.sync()


def nccl_then_adam():
    comm.all_reduce(big, mean=True); comm.sync(); adam_local()


def fused():
    grp.adam(g_off, p_off, m, v, coef, n, 1e-3, 0.9, 0.999, s_adam); grp.sync()


res = {"bucket_mib": n * 4 >> 20, "world": world}
for name, fn in (("nccl_all_reduce_ms", nccl_ar), ("peer_all_reduce_ms", peer_ar), ("nccl_all_reduce_plus_adam_ms", nccl_then_adam),
                 ("fused_rs_adam_ag_ms", fused)):
    res[name] = round(timed(fn), 4)
bytes_wire = 2 * (world - 1) / world * n * 4
res["nccl_busbw_gbs"] = round(bytes_wire / (res["nccl_all_reduce_ms"] * 1e-3) / 1e9, 1)
res["peer_busbw_gbs"] = round(bytes_wire / (res["peer_all_reduce_ms"] * 1e-3) / 1e9, 1)
if rank == 0:
    print(json.dumps(res), flush=True)
dv.sync(); dist.barrier()
grp.close(); comm.close()
dist.destroy_process_group()
