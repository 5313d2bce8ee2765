"""CPU checks of the C-ABI boundary: the shared library loads, exports every symbol
include/burn_b200.h declares, and refuses to compute without a GPU (no fallback)."""
import ctypes as C
import re
from pathlib import Path

import pytest

from burn_b200 import _abi as abi

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "burn_b200.h").read_text()


def declared_symbols():
    names = set(re.findall(r"^(?:int32_t|uint64_t|void \*|void|const char \*)\s*(b200_[a-z0-9_]+)\s*\(", HEADER, re.M))
    return sorted(names)


def test_library_loads_and_reports_abi_version():
    lib = abi.load()
    assert lib.b200_abi_version() == int(re.search(r"#define B200_ABI_VERSION (\d+)", HEADER).group(1))


def test_every_declared_symbol_is_exported_and_bound():
    lib = abi.load()
    names = declared_symbols()
    assert len(names) >= 40
    for name in names:
        assert hasattr(lib, name), f"{name} declared in burn_b200.h but not exported"
        assert name in abi.SIGNATURES, f"{name} has no ctypes signature in burn_b200/_abi.py"
    for name in abi.SIGNATURES:
        assert name in names, f"{name} bound in _abi.py but not declared in the header"


def test_host_layer_symbols_are_exported():
    from burn_b200 import fusion
    host_header = (ROOT / "include" / "burn_b200_host.h").read_text()
    names = set(re.findall(r"^(?:int32_t|b200h_id)\s*(b200h_[a-z0-9_]+)\s*\(", host_header, re.M))
    assert len(names) >= 18
    lib = abi.load()
    for name in names:
        assert hasattr(lib, name), f"{name} declared in burn_b200_host.h but not exported"
        assert name in fusion.HOST_SIGNATURES


def test_opcode_table_matches_header():
    enum_body = HEADER[HEADER.index("B200_OP_MOV = 0"):HEADER.index("B200_OP_COUNT")]
    header_ops = re.findall(r"B200_OP_([A-Z0-9_]+)", enum_body)
    assert header_ops == abi._OPCODES
    assert abi.OP_COUNT == len(header_ops)


def test_struct_layouts():
    assert C.sizeof(abi.Tensor) == 8 + 4 + 4 + 8 * 8 + 8 * 8
    assert C.sizeof(abi.TapeOp) == 8
    assert C.sizeof(abi.Tape) == 32


def test_no_gpu_means_loud_failure_not_fallback():
    lib = abi.load()
    n = C.c_int32(-1)
    st = lib.b200_device_count(C.byref(n))
    if st == 0 and n.value > 0:
        pytest.skip("a GPU is present")
    assert lib.b200_init(0) == abi.ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.b200_last_error()
    with pytest.raises(abi.B200Error):
        from burn_b200 import device
        device.DeviceTensor.empty((4,))


def test_nvrtc_specialised_kernels_compile_without_a_device():
    """jit.cu's generator + NVRTC on the bench chain (elementwise linear / rank-3 forms, row and column
    fuse-on-read reductions), compiled for sm_100a — no GPU is touched, nothing is loaded."""
    import ctypes as C
    from burn_b200 import _abi as abi
    n = C.c_uint64()
    abi.check(abi.load().b200_jit_selftest(C.byref(n)))
    assert n.value > 10_000
