"""GPU: the host fusion layer end to end — lazy ops → fusers → one kernel per block — against the
oracle.  Numeric goldens + block inspection, like crates/burn-backend-tests/tests/fusion/*.rs."""
import numpy as np
import pytest

from burn_b200 import _abi as abi
from burn_b200 import fusion as F
from oracle import oracle
from tests import helpers as H

pytestmark = pytest.mark.gpu


def rnd(shape, seed, lo=-1.0, hi=1.0):
    return np.random.default_rng(seed).uniform(lo, hi, shape).astype(np.float32)


@pytest.fixture()
def st(dev):
    s = F.FusionStream()
    yield s
    s.close()


def test_bench_chain_runs_as_one_kernel(st):
    a, b, c = rnd((512, 1024), 0), rnd((512, 1024), 1), rnd((512, 1024), 2)
    m = a < 0
    ta, tb, tc, tm = st.tensor(a), st.tensor(b), st.tensor(c), st.tensor(m)
    t = ta.mul(tb)
    x = t.add(tc); t.drop()
    g = F.gelu(x); x.drop()
    y = g.mask_fill(tm, 0.0); g.drop()
    got = y.numpy()
    (blk,) = st.blocks()
    assert (blk.kind, blk.n_ops, blk.launches) == (F.BLOCK_ELEMWISE, 8, 1)
    want = oracle.float_mask_fill(oracle.gelu(oracle.float_add(oracle.float_mul(a, b), c)), m, 0.0)
    H.assert_close(got, want, H.REL_ELEMWISE, 0.0, "fused chain")


def test_relu_and_kept_intermediate(st):
    x = rnd((100, 64), 3)
    tx = st.tensor(x)
    mask = tx.lower_equal_elem(0.0)
    y = tx.mask_fill(mask, 0.0)        # relu = lower_equal_elem + mask_fill (activation.rs:37-42)
    H.assert_exact(y.numpy(), oracle.relu(x))
    H.assert_exact(mask.numpy(), x <= 0)   # the kept intermediate was materialised too
    (blk,) = st.blocks()
    assert blk.n_outputs == 2 and blk.launches == 1


def test_layer_norm_denominator_reduce_block(st):
    x = rnd((256, 512), 4)
    tx = st.tensor(x)
    sq = tx.mul(tx)
    m = sq.mean_dim(1); sq.drop()
    e = m.add_scalar(1e-5); m.drop()
    d = e.sqrt(); e.drop()
    got = d.numpy()
    (blk,) = st.blocks()
    assert (blk.kind, blk.n_ops, blk.launches) == (F.BLOCK_REDUCE, 4, 1)
    want = oracle.float_sqrt(oracle.float_add_scalar(oracle.float_mean_dim(oracle.float_mul(x, x), 1), 1e-5))
    H.assert_close(got, want, H.REL_REDUCE, 0.0)


def test_softmax_through_the_stream(st):
    x = rnd((64, 256), 5, -3, 3)
    tx = st.tensor(x)
    mx = tx.max_dim(1)
    sh = tx.sub(mx)
    ex = sh.exp(); sh.drop()
    sm = ex.sum_dim(1)
    y = ex.div(sm); ex.drop(); sm.drop(); mx.drop()
    H.assert_close(y.numpy(), oracle.softmax(x, 1), H.REL_REDUCE, 1e-9)
    (blk,) = st.blocks()                                               # the whole chain: one row-resident kernel
    assert (blk.kind, blk.n_ops, blk.launches) == (F.BLOCK_ROWNORM, 5, 1)
    st.clear_blocks()
    z = F.log_softmax(st.tensor(x), 1)
    H.assert_close(z.numpy(), oracle.log_softmax(x, 1), H.REL_REDUCE, 1e-6)
    assert [(b.kind, b.n_ops, b.launches) for b in st.blocks()] == [(F.BLOCK_ROWNORM, 6, 1)]
    st.clear_blocks()
    w = F.softmax(st.tensor(x), 0)                                     # other axes: the decomposed blocks
    H.assert_close(w.numpy(), oracle.softmax(x, 0), H.REL_REDUCE, 1e-9)
    assert all(b.kind != F.BLOCK_ROWNORM for b in st.blocks())
    assert sum(b.launches for b in st.blocks()) == len(st.blocks())   # one launch per block


def test_layer_norm_through_the_stream(st):
    x = rnd((16, 8, 256), 9, -2, 2)
    g, b = rnd((1, 1, 256), 10, 0.5, 1.5), rnd((1, 1, 256), 11, -0.5, 0.5)
    y = F.layer_norm(st.tensor(x), st.tensor(g), st.tensor(b), 1e-5)
    H.assert_close(y.numpy(), oracle.layer_norm(x, g.reshape(256), b.reshape(256), 1e-5), 1e-5, 1e-6)
    assert [(k.kind, k.n_ops, k.launches) for k in st.blocks()] == [(F.BLOCK_ROWNORM, 9, 1)]


def test_argmax_and_sum_dim_exact(st):
    x = rnd((300, 1000), 6)
    tx = st.tensor(x)
    H.assert_exact(tx.argmax(1).numpy(), oracle.float_argmax(x, 1))
    H.assert_close(tx.sum_dim(0).numpy(), oracle.float_sum_dim(x, 0), H.REL_REDUCE, 1e-6 * 300)


def test_linear_gelu_as_one_matmul_block(st):
    x, w, bias = rnd((256, 128), 7, -0.5, 0.5), rnd((128, 512), 8, -0.5, 0.5), rnd((1, 512), 9)
    tx, tw, tb = st.tensor(x), st.tensor(w), st.tensor(bias)
    h = tx.matmul(tw)
    hb = h.add(tb); h.drop()
    y = F.gelu(hb); hb.drop()
    got = y.numpy()
    (blk,) = st.blocks()
    assert blk.kind == F.BLOCK_MATMUL and blk.n_ops == 7
    want = oracle.gelu(oracle.float_add(oracle.float_matmul(x[None], w[None])[0], bias))
    H.assert_close(got, want, 1e-5, 1e-5, "linear+gelu (3xTF32 GEMM, fused epilogue)")


def test_matmul_on_swap_dims_view(st):
    a, b = rnd((96, 64), 10), rnd((80, 64), 11)
    ta, tb = st.tensor(a), st.tensor(b)
    got = ta.matmul(tb.swap_dims(0, 1)).numpy()      # grad·rhsᵀ shape of work
    H.assert_close(got, oracle.float_matmul(a[None], b.T[None])[0], 1e-5, 1e-5)
