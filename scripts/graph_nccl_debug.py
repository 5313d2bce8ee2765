"""2-rank diagnostic: NCCL all-reduce inside a captured CUDA graph (stage-by-stage progress on stderr)."""
import faulthandler, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.dump_traceback_later(40, exit=True)
import numpy as np, torch, torch.distributed as dist
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
def say(*a):
    print(f"[r{rank} {time.time()%1000:.2f}]", *a, file=sys.stderr, flush=True)
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
from burn_b200 import _abi as abi, device as dv, ops
from burn_b200.device import DeviceTensor
from burn_b200.distributed import Communicator
dv.init(lr)
lib = abi.load()
comm = Communicator(rank, world, device=torch.device("cuda", lr))
say("comm up")
x = DeviceTensor.from_numpy(np.full((1 << 20,), float(rank + 1), np.float32))
buf = DeviceTensor.empty((1 << 20,))
mode = sys.argv[1] if len(sys.argv) > 1 else "fork"
def body():
    y = ops.float_mul_scalar(x, 2.0)
    abi.check(lib.b200_memcpy_d2d(buf.data_ptr(), y.data_ptr(), buf.numel * 4, None))
    comm.all_reduce(buf, mean=True)
    comm.sync()
    z = ops.float_add(buf, x)
    abi.check(lib.b200_memcpy_d2d(buf.data_ptr(), z.data_ptr(), buf.numel * 4, None))
body(); dv.sync(); say("eager ok", buf.numpy()[:2])
say("capture begin")
with dv.Graph.capture() as g:
    body()
say("capture end: nodes", g.total_nodes, "kernels", g.kernel_nodes)
for i in range(3):
    g.launch(); say("launched", i)
    dv.sync(); say("synced", i, buf.numpy()[:2])
dist.barrier(); say("done")
g.destroy(); say('graph destroyed')
comm.close(); say('comm closed')
dist.destroy_process_group()
