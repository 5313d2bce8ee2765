"""Scratch timing of the hot kernels at the BASELINE size (not the contract bench)."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from burn_b200 import _abi as abi, device as dv
from burn_b200.device import DeviceTensor, TapeBuilder
from tests import helpers as H
from tests.test_elemwise_gpu import bench_chain_tape

dv.init(0)
lib = abi.load()
n = int(os.environ.get("N", 8192))
rng = np.random.default_rng(0)
a = rng.uniform(-1, 1, (n, n)).astype(np.float32)
da, db, dc = H.up(a), H.up(rng.uniform(-1, 1, (n, n)).astype(np.float32)), H.up(rng.uniform(-1, 1, (n, n)).astype(np.float32))
dm = H.up(a < 0)
out = DeviceTensor.empty((n, n))

def timeit(fn, iters=20, warm=3):
    for _ in range(warm): fn()
    e0, e1 = C.c_void_p(), C.c_void_p()
    lib.b200_event_create(C.byref(e0)); lib.b200_event_create(C.byref(e1))
    lib.b200_event_record(e0, None)
    for _ in range(iters): fn()
    lib.b200_event_record(e1, None)
    ms = C.c_float()
    abi.check(lib.b200_event_elapsed_ms(e0, e1, C.byref(ms)))
    return ms.value / iters

peak = 6558.7
def report(name, ms, bytes_):
    gbs = bytes_ / ms / 1e6
    print(f"{name:34s} {ms*1e3:9.1f} us  {gbs:8.1f} GB/s  {gbs/peak*100:5.1f}% of measured copy peak", flush=True)

tape = bench_chain_tape().build()
report("chain mask_fill(gelu(a*b+c))", timeit(lambda: dv.launch_elemwise(tape, [da, db, dc, dm], [out], (n, n))), n*n*17)
t2 = TapeBuilder().op("ADD_F", ("in", 0), ("in", 1), out=0).build()
report("add", timeit(lambda: dv.launch_elemwise(t2, [da, db], [out], (n, n))), n*n*12)
t3 = TapeBuilder().op("MUL_F", ("in", 0), ("f", 2.0), out=0).build()
report("mul_scalar", timeit(lambda: dv.launch_elemwise(t3, [da], [out], (n, n))), n*n*8)
t4 = TapeBuilder().op("ERF_F", ("in", 0), out=0).build()
report("erf", timeit(lambda: dv.launch_elemwise(t4, [da], [out], (n, n))), n*n*8)
t5 = TapeBuilder().op("EXP_F", ("in", 0), out=0).build()
report("exp", timeit(lambda: dv.launch_elemwise(t5, [da], [out], (n, n))), n*n*8)
t6 = TapeBuilder().op("TANH_F", ("in", 0), out=0).build()
report("tanh", timeit(lambda: dv.launch_elemwise(t6, [da], [out], (n, n))), n*n*8)
o1 = DeviceTensor.empty((n, 1)); o0 = DeviceTensor.empty((1, n)); oi = DeviceTensor.empty((n, 1), abi.I32); of = DeviceTensor.empty((1,))
oi0 = DeviceTensor.empty((1, n), abi.I32)
report("sum_dim(1)", timeit(lambda: dv.launch_reduce(abi.RED_SUM, 1, (n, n), [da], [o1])), n*n*4 + n*4)
report("sum_dim(0)", timeit(lambda: dv.launch_reduce(abi.RED_SUM, 0, (n, n), [da], [o0])), n*n*4 + n*4)
report("mean_dim(1)", timeit(lambda: dv.launch_reduce(abi.RED_MEAN, 1, (n, n), [da], [o1])), n*n*4 + n*4)
report("argmax(1)", timeit(lambda: dv.launch_reduce(abi.RED_ARGMAX, 1, (n, n), [da], [oi])), n*n*4 + n*4)
report("argmax(0)", timeit(lambda: dv.launch_reduce(abi.RED_ARGMAX, 0, (n, n), [da], [oi0])), n*n*4 + n*4)
report("sum (full)", timeit(lambda: dv.launch_reduce_full(abi.RED_SUM, da, of)), n*n*4)
tb = TapeBuilder().op("MUL_F", ("in", 0), ("in", 1), tmp=0); H.gelu_tape(tb, ("tmp", 0)); tr = tb.build()
report("sum_dim(gelu(a*b),1) fused", timeit(lambda: dv.launch_reduce(abi.RED_SUM, 1, (n, n), [da, db], [o1], read=tr)), n*n*8 + n*4)

# ---- matmul (tensor pipe)
from burn_b200 import ops
import ctypes as C
def mm_report(name, n, prec, bf16=False):
    rng = np.random.default_rng(1)
    a = rng.uniform(-0.5, 0.5, (n, n)).astype(np.float32); b = rng.uniform(-0.5, 0.5, (n, n)).astype(np.float32)
    if bf16:
        da, db = DeviceTensor.from_bf16_of(a), DeviceTensor.from_bf16_of(b).swap_dims(0, 1)
    else:
        da, db = H.up(a), H.up(b).swap_dims(0, 1)      # NT: both operands K-major, no repack
    out = DeviceTensor.empty((n, n))
    ad, bd, cd = da.desc(), db.desc(), out.desc()
    wsb = C.c_uint64(); abi.check(lib.b200_matmul_workspace_bytes(C.byref(ad), C.byref(bd), prec, C.byref(wsb)))
    ws = dv.Storage(max(wsb.value, 16))
    fn = lambda: abi.check(lib.b200_launch_matmul(C.byref(ad), C.byref(bd), C.byref(cd), prec, None, None, 0, ws.ptr, wsb.value, None))
    ms = timeit(fn, iters=10, warm=2)
    tf = 2 * n**3 / ms / 1e9
    print(f"{name:34s} n={n:5d} {ms*1e3:9.1f} us  {tf:8.1f} TFLOP/s  ({tf/1671.5*100:5.1f}% of measured bf16 cuBLAS peak)", flush=True)
for n in (2048, 4096, 8192):
    mm_report("matmul tf32 NT", n, abi.MM_TF32)
    mm_report("matmul bf16 NT (bf16 operands)", n, abi.MM_BF16, bf16=True)
mm_report("matmul f32x3 NT", 4096, abi.MM_F32X3)
