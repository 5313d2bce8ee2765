import os
import sys

import pytest

# a tape that qualifies for NVRTC specialisation must compile: no silent fall back to the interpreter in tests
os.environ.setdefault("B200_TAPE_JIT_STRICT", "1")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with `-m gpu` on the GPU box)")


@pytest.fixture(scope="session")
def dev():
    """Initialises device 0 through the C ABI.  Fails loudly when the CUDA library or a GPU is
    missing — GPU tests never fall back to a CPU path."""
    from burn_b200 import device
    device.init(0)
    return device
