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


# ---------------------------------------------------------------- round 2: relative plans, state, in place, views
def _bias_gelu_softmax(st, x, w, bias, scale):
    tx, tw, tb = st.tensor(x), st.tensor(w), st.tensor(bias)
    h = tx.matmul(tw)
    hb = h.add(tb); h.drop()
    g = F.gelu(hb); hb.drop()
    a = g.mul_scalar(scale); g.drop()
    y = F.softmax(a, 1); a.drop()
    return y


def _bias_gelu_softmax_ref(x, w, bias, scale):
    h = oracle.float_add(oracle.float_matmul(x[None], w[None])[0], bias)
    return oracle.softmax(oracle.float_mul_scalar(oracle.gelu(h), scale), 1)


def test_cached_plan_reexecutes_with_new_shapes_and_scalars(st):
    """A plan built once is re-executed through a new Context (context.rs:11-26): other extents, other scalar."""
    x, w, b = rnd((64, 32), 20, -0.5, 0.5), rnd((32, 128), 21, -0.5, 0.5), rnd((1, 128), 22)
    H.assert_close(_bias_gelu_softmax(st, x, w, b, 0.5).numpy(), _bias_gelu_softmax_ref(x, w, b, 0.5), 2e-5, 1e-7)
    first = [(k.kind, k.n_ops) for k in st.blocks()]
    assert all(k.from_cache == 0 for k in st.blocks())
    st.clear_blocks()
    x, w, b = rnd((200, 96), 23, -0.5, 0.5), rnd((96, 320), 24, -0.5, 0.5), rnd((1, 320), 25)
    H.assert_close(_bias_gelu_softmax(st, x, w, b, 1.75).numpy(), _bias_gelu_softmax_ref(x, w, b, 1.75), 2e-5, 1e-7)
    assert [(k.kind, k.n_ops) for k in st.blocks()] == first
    assert all(k.from_cache == 1 for k in st.blocks())
    cs = st.cache_stats()
    assert cs.hits >= 1 and cs.plans >= 1


def test_optimization_state_round_trip_on_the_device(st):
    def record(x, m, c):
        tx, tm = st.tensor(x), st.tensor(m)
        t = tx.mul_scalar(c)
        u = t.exp(); t.drop()
        return u.mask_fill(tm, -1.0), u
    x1 = rnd((32, 48), 30)
    y1, u1 = record(x1, x1 > 0, 0.5)
    u1.drop()
    opt = st.fuser(F.BLOCK_ELEMWISE).fuse_until_closed().finish()
    state = opt.to_state()
    opt.execute(st)
    H.assert_close(y1.numpy(), oracle.float_mask_fill(oracle.float_exp(oracle.float_mul_scalar(x1, 0.5)), x1 > 0, -1.0), H.REL_ELEMWISE)
    x2 = rnd((100, 20), 31)
    y2, u2 = record(x2, x2 < 0.3, -2.0)
    u2.drop()
    F.Optimization.from_state(state).execute(st)         # rebuilt from bytes, bound to new tensors and scalars
    H.assert_close(y2.numpy(), oracle.float_mask_fill(oracle.float_exp(oracle.float_mul_scalar(x2, -2.0)), x2 < 0.3, -1.0), H.REL_ELEMWISE)


def test_in_place_output_matches_the_out_of_place_result(st):
    """tests/fusion/inplace.rs: an output written over a consumed input must equal the fresh-buffer result."""
    x, o = rnd((257, 129), 40), rnd((257, 129), 41)
    tx, to = st.tensor(x), st.tensor(o)
    y = tx.add(to); tx.drop()
    z = F.gelu(y); y.drop()
    got = z.numpy()
    blk = st.blocks()[0]
    assert blk.aliased == 1 and blk.launches == 1
    H.assert_close(got, oracle.gelu(oracle.float_add(x, o)), H.REL_ELEMWISE)
    H.assert_exact(to.numpy(), o)                        # the other input is untouched


def test_views_in_the_stream(st):
    x = rnd((6, 8, 10), 50)
    tx = st.tensor(x)
    H.assert_close(tx.reshape((48, 10)).exp().numpy(), oracle.float_exp(x.reshape(48, 10)), H.REL_ELEMWISE)
    sl = tx.slice([(1, 5), (2, 8), (0, 10)])
    H.assert_exact(sl.mul_scalar(2.0).numpy(), oracle.float_mul_scalar(x[1:5, 2:8, :], 2.0))
    row = rnd((1, 1, 10), 51)
    ex = st.tensor(row).expand((6, 8, 10))
    H.assert_exact(tx.add(ex).numpy(), oracle.float_add(x, np.broadcast_to(row, x.shape)))
    # a view of a pending tensor: lone view block, then the consumer reads through the new strides
    t = tx.mul_scalar(3.0)
    v = t.swap_dims(0, 2)
    w = v.reshape((10, 48))                              # strided → copies, then reshapes
    H.assert_exact(w.numpy(), np.ascontiguousarray(oracle.float_mul_scalar(x, 3.0).transpose(2, 1, 0)).reshape(10, 48))
    kinds = [b.kind for b in st.blocks()]
    assert F.BLOCK_VIEW in kinds


def test_gather_and_select_through_the_stream(st):
    x = rnd((50, 24), 60)
    rng = np.random.default_rng(61)
    idx = rng.integers(0, 50, size=17).astype(np.int64)
    gi = rng.integers(0, 24, size=(50, 7)).astype(np.int64)
    tx = st.tensor(x)
    s = tx.select(0, st.tensor(idx))
    H.assert_close(s.exp().numpy(), oracle.float_exp(oracle.float_select(x, 0, idx)), H.REL_ELEMWISE)
    g = tx.gather(1, st.tensor(gi))
    H.assert_exact(g.numpy(), oracle.float_gather(1, x, gi))
    assert [b.kind for b in st.blocks()].count(F.BLOCK_EAGER) == 2


def test_wrap_device_memory_and_read_it_back(st, dev):
    from burn_b200.device import DeviceTensor
    a = rnd((33, 65), 70)
    da = DeviceTensor.from_numpy(a)
    la = st.wrap(da.desc())
    y = la.mul_scalar(2.0); la.drop()                    # dropped, but caller-owned memory is never reused in place
    d = y.device_tensor()
    assert d.ptr and d.rank == 2 and st.blocks()[0].aliased == 0
    H.assert_exact(y.numpy(), oracle.float_mul_scalar(a, 2.0))
    H.assert_exact(da.numpy(), a)


def test_encoder_forward_through_the_stream_matches_the_hand_sequenced_forward(st, dev):
    """burn-nn's primitive op stream for the TransformerEncoder (burn_b200/stream_model.py) → fusers → kernels,
    against burn_b200.train's hand-sequenced forward (same GEMM precision, unfused attention chain)."""
    from burn_b200 import train as T
    from burn_b200.device import DeviceTensor
    from burn_b200.stream_model import StreamEncoder
    B, S, d = 4, 48, 128
    model = T.Encoder(5, d, 256, 2, 2)
    x = rnd((B, S, d), 80)
    xd = DeviceTensor.from_numpy(x)
    old = T._ATTN_MODE
    T._ATTN_MODE = "chain"
    try:
        want = model.forward(T.Tape(abi.MM_F32X3), T.Var(xd, False)).v.numpy()
    finally:
        T._ATTN_MODE = old
    enc = StreamEncoder(st, model, abi.MM_F32X3)
    got = enc.forward(st.wrap(xd.desc())).numpy()
    assert np.abs(got - want).max() <= 2e-5 * np.abs(want).max()
    n_first = len(st.blocks())
    assert all(b.launches <= 1 for b in st.blocks() if b.kind != F.BLOCK_MATMUL)   # one kernel per block (a 3xTF32 GEMM splits its operands first)
    st.clear_blocks()
    x2 = rnd((3, 40, d), 81)
    got2 = enc.forward(st.tensor(x2)).numpy()                            # other extents → the cached plan, new Context
    assert len(st.blocks()) == n_first and all(b.from_cache for b in st.blocks())
    T._ATTN_MODE = "chain"
    try:
        want2 = model.forward(T.Tape(abi.MM_F32X3), T.Var(DeviceTensor.from_numpy(x2), False)).v.numpy()
    finally:
        T._ATTN_MODE = old
    assert np.abs(got2 - want2).max() <= 2e-5 * np.abs(want2).max()
