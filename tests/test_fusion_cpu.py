"""CPU: fusion decisions of the host layer (burn_b200/host/fusion.cpp) on plan-only streams —
which IR operations land in which fused block, the analogue of the reference's FusionInspector
tests (crates/burn-backend-tests/tests/fusion/*.rs) and of burn-fusion's fake-backend tests
(crates/burn-fusion/src/stream/execution/tests.rs).  No device is touched."""
import pytest

from burn_b200 import _abi as abi
from burn_b200 import fusion as F


@pytest.fixture()
def st():
    s = F.FusionStream(plan_only=True)
    yield s
    s.close()


def kinds(st):
    return [(b.kind, b.n_ops) for b in st.blocks()]


def test_gelu_chain_with_mask_fill_is_one_elementwise_block(st):
    a, b, c = (st.placeholder((64, 128)) for _ in range(3))
    m = st.placeholder((64, 128), abi.BOOL)
    t = a.mul(b)
    x = t.add(c); t.drop()
    g = F.gelu(x); x.drop()
    y = g.mask_fill(m, 0.0); g.drop()
    st.sync()
    (blk,) = st.blocks()
    assert blk.kind == F.BLOCK_ELEMWISE and blk.n_ops == 8      # mul, add, 5 gelu primitives, mask_fill
    assert blk.n_inputs == 4 and blk.n_outputs == 1             # a, b, c, m in — only y out
    assert blk.n_tape_ops == 8
    assert y.shape == (64, 128)


def test_undropped_intermediates_become_outputs(st):
    a, b = st.placeholder((8, 8)), st.placeholder((8, 8))
    t = a.add(b)          # kept alive by the caller → must be materialised
    u = t.exp()
    st.sync()
    (blk,) = st.blocks()
    assert blk.kind == F.BLOCK_ELEMWISE and blk.n_ops == 2 and blk.n_outputs == 2
    assert u.shape == (8, 8)


def test_shape_change_closes_the_elementwise_block(st):
    a, b = st.placeholder((4, 16)), st.placeholder((4, 16))
    row = st.placeholder((1, 16))
    t = a.add(b)
    u = row.exp()          # different output shape → new block (fuser.rs:774-829 compat rule)
    st.sync()
    assert kinds(st) == [(F.BLOCK_ELEMWISE, 1), (F.BLOCK_ELEMWISE, 1)]
    assert t.shape == (4, 16) and u.shape == (1, 16)


def test_broadcast_inputs_fuse_when_output_shape_is_stable(st):
    x = st.placeholder((32, 64))
    bias = st.placeholder((1, 64))
    t = x.add(bias)
    y = t.mul_scalar(2.0); t.drop()
    st.sync()
    (blk,) = st.blocks()
    assert (blk.kind, blk.n_ops, blk.n_inputs, blk.n_outputs) == (F.BLOCK_ELEMWISE, 2, 2, 1)
    assert y.shape == (32, 64)


def test_reduce_fuser_read_and_write_blocks(st):
    # mean_dim(x*x, 1) then (+eps).sqrt(): fuse-on-read + reduce + fuse-on-write in ONE block
    x = st.placeholder((128, 512))
    sq = x.mul(x)
    m = sq.mean_dim(1); sq.drop()
    e = m.add_scalar(1e-5); m.drop()
    d = e.sqrt(); e.drop()
    st.sync()
    (blk,) = st.blocks()
    assert blk.kind == F.BLOCK_REDUCE and blk.n_ops == 4
    assert blk.n_inputs == 1 and blk.n_outputs == 1
    assert d.shape == (128, 1)


def test_reduce_without_fusable_neighbours(st):
    x = st.placeholder((16, 32))
    s = x.sum_dim(0)
    am = x.argmax(1)
    st.sync()
    assert kinds(st) == [(F.BLOCK_REDUCE, 1), (F.BLOCK_REDUCE, 1)]
    assert s.shape == (1, 32) and am.shape == (16, 1) and am.dtype == abi.I32


def test_read_block_value_still_alive_prevents_fuse_on_read(st):
    x = st.placeholder((16, 32))
    e = x.exp()            # caller keeps `e` → it must be written, so it cannot live in the reduce's read block
    s = e.sum_dim(1)
    st.sync()
    assert kinds(st) == [(F.BLOCK_ELEMWISE, 1), (F.BLOCK_REDUCE, 1)]
    assert s.shape == (16, 1)


def test_softmax_chain_along_the_last_axis_is_one_row_resident_block(st):
    # softmax = max_dim, sub, exp, sum_dim, div (activation.rs:250-256) → the ReduceBroadcasted shape
    x = st.placeholder((64, 256))
    y = F.softmax(x, 1)
    st.sync()
    (blk,) = st.blocks()
    assert blk.kind == F.BLOCK_ROWNORM and blk.n_ops == 5 and blk.n_inputs == 1 and blk.n_outputs == 1
    assert y.shape == (64, 256)
    st.clear_blocks()
    z = F.log_softmax(x, 1)                              # max_dim, sub, exp, sum_dim, log, sub
    st.sync()
    (blk,) = st.blocks()
    assert blk.kind == F.BLOCK_ROWNORM and blk.n_ops == 6
    assert z.shape == (64, 256)


def test_layer_norm_chain_is_one_row_resident_block(st):
    # mean_dim, sub, mul, mean_dim, add_scalar, sqrt, div, mul(gamma), add(beta)  (modules/base.rs:846-877)
    x = st.placeholder((32, 8, 512))
    gamma, beta = st.placeholder((1, 1, 512)), st.placeholder((1, 1, 512))
    y = F.layer_norm(x, gamma, beta, 1e-5)
    st.sync()
    (blk,) = st.blocks()
    assert blk.kind == F.BLOCK_ROWNORM and blk.n_ops == 9 and blk.n_inputs == 3 and blk.n_outputs == 1
    assert y.shape == (32, 8, 512)
    st.clear_blocks()
    z = F.layer_norm(x, gamma, None, 1e-5)
    st.sync()
    (blk,) = st.blocks()
    assert blk.kind == F.BLOCK_ROWNORM and blk.n_ops == 8 and blk.n_inputs == 2
    assert z.shape == (32, 8, 512)


def test_softmax_with_a_live_intermediate_or_another_axis_decomposes(st):
    x = st.placeholder((64, 256))
    mx = x.max_dim(1)
    sh = x.sub(mx)
    ex = sh.exp(); sh.drop()
    sm = ex.sum_dim(1)
    y = ex.div(sm); ex.drop(); sm.drop()                 # mx stays alive: it must be materialised
    st.sync()
    ks = kinds(st)
    assert ks[0] == (F.BLOCK_REDUCE, 1)                  # max_dim on its own
    assert all(k != F.BLOCK_ROWNORM for k, _ in ks) and sum(n for _, n in ks) == 5
    assert y.shape == (64, 256) and mx.shape == (64, 1)
    st.clear_blocks()
    w = F.softmax(x, 0)                                  # not the last axis
    st.sync()
    assert all(k != F.BLOCK_ROWNORM for k, _ in kinds(st)) and sum(n for _, n in kinds(st)) == 5
    assert w.shape == (64, 256)


def test_matmul_with_bias_gelu_epilogue_is_one_block(st):
    x, w = st.placeholder((256, 128)), st.placeholder((128, 512))
    bias = st.placeholder((1, 512))
    h = x.matmul(w)
    hb = h.add(bias); h.drop()
    y = F.gelu(hb); hb.drop()
    st.sync()
    (blk,) = st.blocks()
    assert blk.kind == F.BLOCK_MATMUL and blk.n_ops == 7     # matmul + add + 5 gelu primitives
    assert blk.n_inputs == 3 and blk.n_outputs == 1
    assert y.shape == (256, 512)


def test_matmul_result_kept_alive_disables_the_epilogue(st):
    x, w = st.placeholder((64, 64)), st.placeholder((64, 64))
    h = x.matmul(w)        # `h` not dropped → raw product must be written, epilogue runs separately
    y = h.exp()
    st.sync()
    assert kinds(st) == [(F.BLOCK_MATMUL, 1), (F.BLOCK_ELEMWISE, 1)]
    assert y.shape == (64, 64)


def test_long_chain_is_split_at_the_tape_limit(st):
    x = st.placeholder((8, 8))
    cur = x
    for i in range(100):
        nxt = cur.add_scalar(1.0)
        if cur is not x:
            cur.drop()
        cur = nxt
    st.sync()
    ks = kinds(st)
    assert all(k == F.BLOCK_ELEMWISE for k, _ in ks)
    assert [n for _, n in ks] == [64, 36]               # B200_MAX_TAPE_OPS = 64 (fuser.rs:857)


def test_shape_errors_surface_at_registration(st):
    a, b = st.placeholder((4, 5)), st.placeholder((4, 6))
    with pytest.raises(abi.B200Error):
        a.add(b)
    with pytest.raises(abi.B200Error):
        a.matmul(b)
    with pytest.raises(abi.B200Error):
        a.sum_dim(2)
