"""GEMM throughput + spot-check sweep (configs[2] shapes).  usage: gemm_sweep.py [quick]"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from burn_b200 import _abi as abi, device as dv
from burn_b200.device import DeviceTensor
from tests import helpers as H
dv.init(0); lib = abi.load()
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"

def timed(fn, iters):
    e0, e1 = C.c_void_p(), C.c_void_p()
    abi.check(lib.b200_event_create(C.byref(e0))); abi.check(lib.b200_event_create(C.byref(e1)))
    for _ in range(3): fn()
    dv.sync()
    abi.check(lib.b200_event_record(e0, None))
    for _ in range(iters): fn()
    abi.check(lib.b200_event_record(e1, None))
    ms = C.c_float(); abi.check(lib.b200_event_elapsed_ms(e0, e1, C.byref(ms)))
    return ms.value / iters

def timed_graph(fn, iters):
    """The same launches replayed from a CUDA graph: no per-launch host work (ctypes, descriptor encode) in the
    timed region — what a captured training step sees."""
    from burn_b200.device import Graph
    with Graph.capture() as g:
        for _ in range(iters): fn()
    g.launch(); dv.sync()
    e0, e1 = C.c_void_p(), C.c_void_p()
    abi.check(lib.b200_event_create(C.byref(e0))); abi.check(lib.b200_event_create(C.byref(e1)))
    abi.check(lib.b200_event_record(e0, None))
    g.launch()
    abi.check(lib.b200_event_record(e1, None))
    dv.sync()
    ms = C.c_float(); abi.check(lib.b200_event_elapsed_ms(e0, e1, C.byref(ms)))
    g.destroy()
    return ms.value / iters

def case(batch, m, n, k, prec, layout="NT", check=True):
    rng = np.random.default_rng(m + n + k)
    bs = (batch,) if batch > 1 else ()
    a = rng.uniform(-0.5, 0.5, bs + (m, k)).astype(np.float32)
    b = rng.uniform(-0.5, 0.5, bs + (k, n)).astype(np.float32)
    bf = prec == abi.MM_BF16
    mk = (lambda x: DeviceTensor.from_bf16_of(x)) if bf else H.up
    nd = len(bs) + 2
    da = mk(a) if layout[0] == "N" else mk(np.ascontiguousarray(np.swapaxes(a, -1, -2))).swap_dims(nd - 2, nd - 1)
    db = mk(b) if layout[1] == "N" else mk(np.ascontiguousarray(np.swapaxes(b, -1, -2))).swap_dims(nd - 2, nd - 1)
    out = DeviceTensor.empty(bs + (m, n))
    ad, bd, cd = da.desc(), db.desc(), out.desc()
    wsb = C.c_uint64()
    abi.check(lib.b200_matmul_workspace_bytes(C.byref(ad), C.byref(bd), prec, C.byref(wsb)))
    ws = dv.Storage(wsb.value) if wsb.value else None
    fn = lambda: abi.check(lib.b200_launch_matmul(C.byref(ad), C.byref(bd), C.byref(cd), prec, None, None, 0,
                                                  ws.ptr if ws else None, wsb.value, None))
    ms = timed(fn, 5 if quick else 20)
    msg = timed_graph(fn, 20) if m * n * k <= 4096 ** 3 else ms
    err = -1.0
    if check:
        got = out.numpy()
        rows = rng.integers(0, m, 16)
        sl = (0,) * len(bs)
        ref = a[sl][rows].astype(np.float64) @ b[sl].astype(np.float64)
        bound = np.abs(a[sl][rows]).astype(np.float64) @ np.abs(b[sl]).astype(np.float64)
        err = float(np.max(np.abs(got[sl][rows] - ref) / bound))
        sl2 = (batch - 1,) * len(bs)
        ref2 = a[sl2][rows].astype(np.float64) @ b[sl2].astype(np.float64)
        err = max(err, float(np.max(np.abs(got[sl2][rows] - ref2) / (np.abs(a[sl2][rows]).astype(np.float64) @ np.abs(b[sl2]).astype(np.float64)))))
    tf = 2.0 * batch * m * n * k / (ms * 1e-3) / 1e12
    name = {abi.MM_TF32: "tf32", abi.MM_BF16: "bf16", abi.MM_F32X3: "f32x3"}[prec]
    tfg = 2.0 * batch * m * n * k / (msg * 1e-3) / 1e12
    print(f"{name:5s} {layout} b{batch:<3d} {m:6d}x{n:6d}x{k:6d}  {ms:9.4f} ms  {tf:8.1f} TF/s  graph-replay {msg:9.4f} ms {tfg:8.1f} TF/s  relerr {err:.2e}", flush=True)

def epi_case(m, n, k, prec):
    """configs[2]: fused bias + GELU epilogue, timed against the plain GEMM of the same shape."""
    from burn_b200.device import TapeBuilder
    rng = np.random.default_rng(7)
    a = rng.uniform(-0.5, 0.5, (m, k)).astype(np.float32); b = rng.uniform(-0.5, 0.5, (n, k)).astype(np.float32)
    bias = rng.uniform(-0.5, 0.5, (1, n)).astype(np.float32)
    bf = prec == abi.MM_BF16
    mk = (lambda x: DeviceTensor.from_bf16_of(x)) if bf else H.up
    da, db, dbias = mk(a), mk(b).swap_dims(0, 1), H.up(bias)
    out = DeviceTensor.empty((m, n))
    tb = TapeBuilder().op("ADD_F", ("in", 0), ("in", 1), tmp=0)
    H.gelu_tape(tb, ("tmp", 0), out=0)
    tape = tb.build()
    ad, bd, cd, ed = da.desc(), db.desc(), out.desc(), dbias.desc()
    fn = lambda: abi.check(lib.b200_launch_matmul(C.byref(ad), C.byref(bd), C.byref(cd), prec, C.byref(tape), C.byref(ed), 1, None, 0, None))
    ms = timed(fn, 5 if quick else 20)
    name = {abi.MM_TF32: "tf32", abi.MM_BF16: "bf16"}[prec]
    print(f"{name:5s} NT +bias+gelu epilogue {m:6d}x{n:6d}x{k:6d}  {ms:9.4f} ms  {2.0*m*n*k/(ms*1e-3)/1e12:8.1f} TF/s", flush=True)

shapes = [(1, 256, 256, 256), (1, 384, 640, 200), (1, 1024, 1024, 1024), (1, 4096, 4096, 4096), (1, 8192, 8192, 8192)]
if not quick:
    shapes += [(1, 2048, 2048, 2048), (1, 16384, 16384, 16384), (64, 2048, 2048, 2048), (1, 8192, 1024, 1024), (1, 8192, 4096, 1024), (1, 8192, 1024, 4096), (1, 1024, 4096, 8192)]
for prec in (abi.MM_BF16, abi.MM_TF32):
    for (bt, m, n, k) in shapes:
        if m * k * 4 > 2.2e9: check = False   # 16384^3 is asserted in tests/test_matmul_gpu.py instead
        else: check = True
        case(bt, m, n, k, prec, "NT", check)
for lay in ("NN", "TN", "TT"):
    case(1, 4096, 4096, 4096, abi.MM_TF32, lay)
    case(1, 4096, 4096, 4096, abi.MM_BF16, lay)
for prec in (abi.MM_BF16, abi.MM_TF32):
    epi_case(4096, 4096, 4096, prec)
    if not quick:
        epi_case(8192, 8192, 8192, prec)
if not quick:
    case(1, 4096, 4096, 4096, abi.MM_F32X3, "NN")
    case(1, 8192, 8192, 8192, abi.MM_F32X3, "NN", check=False)
