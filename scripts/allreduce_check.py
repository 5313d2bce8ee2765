"""Multi-rank check of b200_all_reduce (run under torchrun): Sum and Mean against the closed
form, then bus bandwidth on a 256 MiB f32 gradient bucket — the DDP exchange step of SURVEY §8(e)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from burn_b200 import _abi as abi, device as dv, ops
from burn_b200.device import DeviceTensor
from burn_b200.distributed import Communicator
dv.init(local)
lib = abi.load()
comm = Communicator(rank, world, device=torch.device("cuda", local))
n = 1 << 20
x = np.arange(n, dtype=np.float32) % 97 + rank
t = DeviceTensor.from_numpy(x)
comm.all_reduce(t, mean=False); comm.sync(); dv.sync()
want = (np.arange(n, dtype=np.float32) % 97) * world + sum(range(world))
assert np.array_equal(t.numpy(), want), "Sum mismatch"
t = DeviceTensor.from_numpy(x)
comm.all_reduce(t, mean=True); comm.sync(); dv.sync()
assert np.allclose(t.numpy(), want / world, rtol=1e-6), "Mean mismatch"
# bandwidth
big = DeviceTensor.empty((64 << 20,))
abi.check(lib.b200_memset(big.data_ptr(), 0, big.numel * 4, None))
for _ in range(3):
    comm.all_reduce(big, mean=True)
comm.sync(); dv.sync(); dist.barrier()
e0, e1 = C.c_void_p(), C.c_void_p()
lib.b200_event_create(C.byref(e0)); lib.b200_event_create(C.byref(e1))
lib.b200_event_record(e0, None)
iters = 10
for _ in range(iters):
    comm.all_reduce(big, mean=True)
comm.sync()
lib.b200_event_record(e1, None)
ms = C.c_float(); abi.check(lib.b200_event_elapsed_ms(e0, e1, C.byref(ms)))
tt = torch.tensor([ms.value], device="cuda", dtype=torch.float64); dist.all_reduce(tt, op=dist.ReduceOp.MAX)
if rank == 0:
    sec = tt.item() / 1e3 / iters
    bytes_ = big.numel * 4
    print(f"all_reduce ok on {world} ranks; 256 MiB f32: {sec*1e3:.3f} ms, algbw {bytes_/sec/1e9:.1f} GB/s, busbw {bytes_/sec/1e9*2*(world-1)/world:.1f} GB/s")
comm.close()
dist.barrier(); dist.destroy_process_group()
