"""GPU: NVRTC-specialised elementwise kernels (burn_b200/csrc/jit.cu) against the op-tape interpreter.
Large linear launches are compiled from the tape; the same launch with B200_TAPE_JIT_MIN_VEC raised
runs on the interpreter.  Both must agree bit for bit (same eval code, opcode folded at compile time)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r'''
import sys, hashlib
import numpy as np
sys.path.insert(0, %r)
from burn_b200 import _abi as abi, device as dv, ops
from burn_b200.device import DeviceTensor, TapeBuilder
from tests import helpers as H
dv.init(0)
n = (1 << 20) + 4 * 37          # > the JIT threshold, ragged against the 512-vector block
rng = np.random.default_rng(0)
a = rng.uniform(-3, 3, n).astype(np.float32); b = rng.uniform(-3, 3, n).astype(np.float32)
c = rng.uniform(0.1, 3, n).astype(np.float32); m = rng.random(n) < 0.3
ia = rng.integers(-100, 100, n).astype(np.int32); ib = rng.integers(1, 17, n).astype(np.int32)
a[:8] = [0.0, -0.0, np.inf, -np.inf, np.nan, 1e-30, -1e30, 4.5]
da, db, dc, dm, dia, dib = (H.up(x) for x in (a, b, c, m, ia, ib))
case = [0]
def run(tb, ins, n_out=1, dts=(abi.F32,)):
    outs = [DeviceTensor.empty((n,), dts[i]) for i in range(n_out)]
    tape = tb.build()
    dv.launch_elemwise(tape, ins, outs, (n,))
    h = hashlib.sha256()
    for o in outs:
        h.update(o.numpy().tobytes())
    names = "+".join(k for o in tb.ops for k, v in abi.OP.items() if v == o[0])
    print("CASE", case[0], names, h.hexdigest()[:16])
    case[0] += 1
# the bench chain
tb = TapeBuilder(); tb.op("MUL_F", ("in", 0), ("in", 1)); tb.op("ADD_F", "acc", ("in", 2), tmp=0)
H.gelu_tape(tb, ("tmp", 0)); tb.op("SELECT", "acc", ("f", 0.0), ("in", 3), out=0)
run(tb, [da, db, dc, dm])
# every float unary / binary op, comparisons, int ops, casts, scalar division, two outputs
for name in ["NEG_F", "ABS_F", "EXP_F", "LOG_F", "LOG1P_F", "SQRT_F", "RECIP_F", "TANH_F", "ERF_F", "FLOOR_F", "CEIL_F",
             "ROUND_F", "TRUNC_F", "SIGN_F", "SIGMOID_F", "SIN_F", "COS_F", "ATAN_F"]:
    run(TapeBuilder().op(name, ("in", 0), out=0), [dc if name in ("LOG_F", "SQRT_F") else da])
for name in ["ADD_F", "SUB_F", "MUL_F", "DIV_F", "REM_F", "REMT_F", "POW_F", "MIN_F", "MAX_F", "ATAN2_F"]:
    run(TapeBuilder().op(name, ("in", 0), ("in", 1), out=0), [dc if name == "POW_F" else da, db])
for name in ["EQ_F", "NE_F", "LT_F", "LE_F", "GT_F", "GE_F"]:
    run(TapeBuilder().op(name, ("in", 0), ("in", 1), out=0), [da, db], dts=(abi.BOOL,))
for name in ["ADD_I", "SUB_I", "MUL_I", "DIV_I", "REM_I", "MIN_I", "MAX_I", "AND_I", "OR_I", "XOR_I", "SHL_I", "SHR_I"]:
    run(TapeBuilder().op(name, ("in", 0), ("in", 1), out=0), [dia, dib], dts=(abi.I32,))
run(TapeBuilder().op("DIV_F", ("in", 0), ("f", 3.0), out=0), [da])
run(TapeBuilder().op("DIV_F", ("in", 0), ("f", 8.0), out=0), [da])
run(TapeBuilder().op("CLAMP_F", ("in", 0), ("f", -1.0), ("f", 1.5), out=0), [da])
run(TapeBuilder().op("F2I", ("in", 0), out=0), [db], dts=(abi.I32,))
run(TapeBuilder().op("I2F", ("in", 0), out=0), [dia])
run(TapeBuilder().op("B2F", ("in", 0), out=0), [dm])
tb = TapeBuilder(); tb.op("MUL_F", ("in", 0), ("in", 1), tmp=0, out=1); tb.op("ADD_F", ("tmp", 0), ("in", 2)); tb.op("EXP_F", "acc", out=0)
run(tb, [da, db, H.up(np.float32([0.25]).reshape(1)).expand((n,))], n_out=2, dts=(abi.F32, abi.F32))
# half-precision storage on both sides of a specialised tape: bf16 / f16 inputs and OUTPUTS (rounded RN-even by the same
# instructions in both modes)
hb = H.up((rng.uniform(-3, 3, n).astype(np.float32).view(np.uint32) >> 16).astype(np.uint16), abi.BF16)
hf = H.up(rng.uniform(-3, 3, n).astype(np.float16).view(np.uint16), abi.F16)
for odt in (abi.BF16, abi.F16):
    tb = TapeBuilder(); tb.op("MUL_F", ("in", 0), ("in", 1)); tb.op("ADD_F", "acc", ("in", 2)); tb.op("TANH_F", "acc", out=0)
    run(tb, [hb, hf, da], dts=(odt,))
    tb = TapeBuilder(); tb.op("MUL_F", ("in", 0), ("f", 1.0009765625), out=0); tb.op("EXP_F", "acc", out=1)
    run(tb, [hf if odt == abi.F16 else hb], n_out=2, dts=(odt, abi.F32))
# rank-3 specialisation: row / column broadcasts, a sliced (row-strided) operand and a strided output
R, Cc = 2048, 1024
x2 = rng.uniform(-2, 2, (R, Cc)).astype(np.float32); rowv = rng.uniform(-2, 2, (1, Cc)).astype(np.float32)
colv = rng.uniform(0.5, 2, (R, 1)).astype(np.float32); wide = rng.uniform(-2, 2, (R, 2 * Cc)).astype(np.float32)
m2 = rng.random((1, Cc)) < 0.4
dx2, drow, dcol, dwide, dm2 = (H.up(t) for t in (x2, rowv, colv, wide, m2))
def run2(tb, ins, out=None):
    out = out if out is not None else DeviceTensor.empty((R, Cc))
    dv.launch_elemwise(tb.build(), ins, [out], (R, Cc))
    names = "+".join(k for o in tb.ops for k, v in abi.OP.items() if v == o[0])
    print("CASE", case[0], "r3:" + names, hashlib.sha256(out.numpy().tobytes()).hexdigest()[:16])
    case[0] += 1
tb = TapeBuilder().op("ADD_F", ("in", 0), ("in", 1)); tb.op("DIV_F", "acc", ("in", 2)); tb.op("SELECT", "acc", ("f", -1.0), ("in", 3), out=0)
run2(tb, [dx2, drow.expand((R, Cc)), dcol.expand((R, Cc)), dm2.expand((R, Cc))])
tb = TapeBuilder().op("MUL_F", ("in", 0), ("in", 1), tmp=0); H.gelu_tape(tb, ("tmp", 0), out=0)
run2(tb, [dwide.slice([(0, R), (Cc, 2 * Cc)]), dx2])
big = DeviceTensor.empty((R, 2 * Cc))
run2(TapeBuilder().op("SUB_F", ("in", 0), ("in", 1), out=0), [dx2, dcol.expand((R, Cc))], out=big.slice([(0, R), (0, Cc)]))
print("DONE")
'''


def _run(env_extra):
    env = dict(os.environ, **env_extra)
    r = subprocess.run([sys.executable, "-c", SCRIPT % ROOT], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "DONE" in r.stdout
    cases = {tuple(l.split()[1:3]): l.split()[3] for l in r.stdout.splitlines() if l.startswith("CASE")}
    return cases, r.stderr


def test_specialised_kernels_match_the_interpreter_bit_for_bit(dev):
    jit, err = _run({"B200_TAPE_JIT": "1"})
    assert "NVRTC failed" not in err and "libnvrtc not found" not in err, err[-2000:]
    interp, _ = _run({"B200_TAPE_JIT": "0"})
    assert len(jit) == len(interp) > 50
    differing = [k for k in interp if jit.get(k) != interp[k]]
    assert not differing, f"specialised kernel != interpreter for {differing}"


EVICT = r'''
import sys, ctypes as C
import numpy as np
sys.path.insert(0, %r)
from burn_b200 import _abi as abi, device as dv
from burn_b200.device import DeviceTensor, TapeBuilder
from oracle import oracle
dv.init(0)
lib = abi.load()
rng = np.random.default_rng(1)
def once(rows):
    x = rng.uniform(-2, 2, (rows, 1024)).astype(np.float32)
    row = rng.uniform(-2, 2, (1, 1024)).astype(np.float32)
    out = DeviceTensor.empty((rows, 1024))
    tb = TapeBuilder().op("ADD_F", ("in", 0), ("in", 1)); tb.op("MUL_F", "acc", ("f", 0.5), out=0)
    dv.launch_elemwise(tb.build(), [DeviceTensor.from_numpy(x), DeviceTensor.from_numpy(row).expand((rows, 1024))], [out], (rows, 1024))
    assert np.array_equal(out.numpy(), oracle.float_mul_scalar(oracle.float_add(x, row), 0.5)), rows
for rows in (1100, 1200, 1300, 1400, 1100, 1500, 1200):      # shape-specialised: every new row count is a new kernel
    once(rows)
n, ev = C.c_uint64(), C.c_uint64()
abi.check(lib.b200_jit_cache_stats(C.byref(n), C.byref(ev)))
print("STATS", n.value, ev.value)
'''


def test_specialised_kernel_cache_is_bounded(dev):
    """Dynamic shapes must not grow the cache (and the loaded cubins) without bound: with room for two kernels, seven
    launches over five shapes evict at least three times and every result is still right."""
    env = dict(os.environ, B200_JIT_CACHE_MAX="2", B200_TAPE_JIT="1", B200_TAPE_JIT_STRICT="1")
    r = subprocess.run([sys.executable, "-c", EVICT % ROOT], capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    entries, evictions = map(int, [l for l in r.stdout.splitlines() if l.startswith("STATS")][0].split()[1:])
    assert entries <= 2 and evictions >= 3
