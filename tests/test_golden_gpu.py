"""GPU: the reference's golden vectors run through the CUDA library (C ABI) — the same fixtures
tests/test_oracle_golden.py pins the oracle with, with the tolerance each reference test uses."""
import pytest

from tests import golden_expr as X
from tests import golden_runner as G

pytestmark = pytest.mark.gpu
CASES = G.load_cases()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_device_reproduces_reference_golden(dev, case):
    G.check(case, G.run_device(case))


EXPR_CASES = X.load_cases()


@pytest.fixture(scope="module")
def device_backend(dev):
    return X.DeviceBackend()


@pytest.mark.parametrize("case", EXPR_CASES, ids=[c["name"] for c in EXPR_CASES])
def test_device_reproduces_reference_test_expression(device_backend, case):
    """The same expression fixtures that pin the oracle, evaluated by the CUDA library through the C ABI."""
    X.check(case, X.evaluate(case["expr"], device_backend))
