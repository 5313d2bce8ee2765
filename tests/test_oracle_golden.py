"""CPU: pins the oracle (oracle/) against the reference's own golden vectors.

Every fixture in tests/golden/burn_backend_tests.json is a literal expected value taken from
crates/burn-backend-tests/tests/tensor/float/** (file:line cited per case).  The oracle must
reproduce all of them within the tolerance the reference test itself uses — that is what lets
the GPU parity tests trust it.
"""
import numpy as np
import pytest

from tests import golden_expr as X
from tests import golden_runner as G

CASES = G.load_cases()
EXPR_CASES = X.load_cases()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_reproduces_reference_golden(case):
    G.check(case, G.run_oracle(case))


@pytest.fixture(scope="module")
def oracle_backend():
    return X.OracleBackend()


@pytest.mark.parametrize("case", EXPR_CASES, ids=[c["name"] for c in EXPR_CASES])
def test_oracle_reproduces_reference_test_expression(case, oracle_backend):
    """tests/golden/burn_backend_tests_expr.json: assertions transliterated from the reference's own tests by
    scripts/extract_goldens.py (op tree over literal tensors → the literal the reference asserts)."""
    X.check(case, X.evaluate(case["expr"], oracle_backend))


def test_expression_fixture_file_is_well_formed():
    assert len(EXPR_CASES) + len(CASES) >= 250
    files = {c["cite"].split(":")[0] for c in EXPR_CASES}
    for must in ("maxmin.rs", "comparison.rs", "mask.rs", "aggregation.rs", "matmul.rs", "reduce_broadcasted.rs"):
        assert any(f.endswith(must) for f in files), must
    for c in EXPR_CASES:
        assert {"name", "cite", "expr", "expected", "tol"} <= set(c)


def test_fixture_file_is_well_formed():
    assert len(CASES) >= 60
    for c in CASES:
        assert {"name", "cite", "op", "inputs", "expected", "tol"} <= set(c)
        assert ":" in c["cite"] and c["cite"].split(":")[1].isdigit()


def test_unrolled_sum_matches_ndarray_order():
    """ndarray::numeric_util::unrolled_fold: 8 partial sums, (p0+p4)+(p1+p5)+(p2+p6)+(p3+p7), tail."""
    from oracle import oracle
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, 1003).astype(np.float32)
    p = [np.float32(0)] * 8
    n8 = (len(x) // 8) * 8
    for i in range(0, n8, 8):
        for k in range(8):
            p[k] = np.float32(p[k] + x[i + k])
    acc = np.float32(0)
    for k in range(4):
        acc = np.float32(acc + np.float32(p[k] + p[k + 4]))
    for v in x[n8:]:
        acc = np.float32(acc + v)
    assert oracle.float_sum(x)[0] == acc


def test_sum_axis_orders():
    from oracle import oracle
    rng = np.random.default_rng(1)
    x = rng.uniform(-1, 1, (7, 19)).astype(np.float32)
    # non-minimal-stride axis: res = res + subview, row after row
    acc = np.zeros(19, dtype=np.float32)
    for r in range(7):
        acc = (acc + x[r]).astype(np.float32)
    assert np.array_equal(oracle.float_sum_dim(x, 0)[0], acc)
    assert oracle.float_sum_dim(x, 1).shape == (7, 1)
    assert np.array_equal(oracle.float_mean_dim(x, 0), (oracle.float_sum_dim(x, 0) / np.float32(7)).astype(np.float32))


def test_argmax_semantics():
    from oracle import oracle
    x = np.array([[1.0, 5.0, 5.0, np.nan, 7.0, np.nan]], dtype=np.float32)
    assert oracle.float_argmax(x, 1).tolist() == [[3]]
    assert oracle.float_argmin(x, 1).tolist() == [[3]]
    assert oracle.float_argmax(np.array([[2.0, 9.0, 9.0]], dtype=np.float32), 1).tolist() == [[1]]
    with pytest.raises(IndexError):
        oracle.float_argmax(x, 2)


def test_erf_tanh_are_f64_then_rounded():
    from oracle import oracle
    import math
    x = np.linspace(-4, 4, 1001).astype(np.float32)
    want = np.array([np.float32(math.erf(float(v))) for v in x], dtype=np.float32)
    assert np.array_equal(oracle.float_erf(x), want)
    want = np.array([np.float32(math.tanh(float(v))) for v in x], dtype=np.float32)
    assert np.array_equal(oracle.float_tanh(x), want)


def test_matmul_broadcast_and_errors():
    from oracle import oracle
    rng = np.random.default_rng(2)
    a = rng.uniform(-1, 1, (2, 1, 5, 7)).astype(np.float32)
    b = rng.uniform(-1, 1, (1, 3, 7, 4)).astype(np.float32)
    got = oracle.float_matmul(a, b)
    assert got.shape == (2, 3, 5, 4)
    ref = np.matmul(a.astype(np.float64), b.astype(np.float64))
    assert np.allclose(got, ref, rtol=1e-5, atol=1e-5)
    with pytest.raises(ValueError):
        oracle.float_matmul(a, rng.uniform(-1, 1, (1, 3, 6, 4)).astype(np.float32))


def test_bench_chain_unfused_equals_op_by_op():
    from oracle import oracle
    rng = np.random.default_rng(3)
    a, b, c = (rng.uniform(-1, 1, (64, 96)).astype(np.float32) for _ in range(3))
    m = a < 0
    want = oracle.float_mask_fill(oracle.gelu(oracle.float_add(oracle.float_mul(a, b), c)), m, 0.0)
    assert np.array_equal(oracle.bench_chain_unfused(a, b, c, m), want)
