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


# ---------------------------------------------------------------- the OperationFuser / Optimization contract
def test_fuser_state_machine_score_and_len(st):
    # two unary ops, the intermediate dropped: 2 reads + 2 writes unfused, 1 + 1 fused, one launch saved
    # → (4 - 2) * 100 + (2 - 1) * 10 = 210, the reference's own scoring test (scoring.rs:115-124)
    x = st.placeholder((32, 32))
    t = x.exp()
    y = t.log(); t.drop()
    f = st.fuser(F.BLOCK_ELEMWISE)
    assert f.status == F.FUSER_OPEN and len(f) == 0 and f.properties() == (0, False)
    f.fuse_next()                                        # exp
    assert len(f) == 1 and f.properties() == (0, True)   # one op: nothing saved yet
    f.fuse_next()                                        # log — `t` not dropped yet, so it is still an output
    assert len(f) == 2 and f.properties() == (110, True)  # 4 unfused − 3 fused (x in, t and y out) + 1 launch
    f.fuse_next()                                        # drop(t): the intermediate stays in registers
    assert f.properties() == (210, True)
    f.reset()
    assert len(f) == 0 and f.status == F.FUSER_OPEN and f.properties() == (0, False)
    st.sync()
    assert y.shape == (32, 32)


def test_fusers_close_on_what_they_cannot_take_and_the_best_score_wins(st):
    x = st.placeholder((64, 128))
    sq = x.mul(x)
    m = sq.mean_dim(1); sq.drop()
    e = m.add_scalar(1e-5); m.drop()
    ew, rd, mm = st.fuser(F.BLOCK_ELEMWISE), st.fuser(F.BLOCK_REDUCE), st.fuser(F.BLOCK_MATMUL)
    for f in (ew, rd, mm):
        f.fuse_until_closed()
    assert mm.status == F.FUSER_CLOSED and len(mm) == 0 and not mm.properties()[1]
    assert ew.status == F.FUSER_CLOSED and len(ew) == 1        # stops at the reduce
    assert len(rd) == 3 and rd.properties()[1]
    assert rd.properties()[0] > ew.properties()[0]
    st.sync()
    (blk,) = st.blocks()
    assert blk.kind == F.BLOCK_REDUCE and blk.score == rd.properties()[0]
    assert e.shape == (64, 1)


def test_clone_dyn_explores_independently(st):
    a, b = st.placeholder((8, 8)), st.placeholder((8, 8))
    t = a.add(b)
    u = t.exp(); t.drop()
    v = u.mul(a); u.drop()
    f = st.fuser(F.BLOCK_ELEMWISE)
    f.fuse_next()
    g = f.clone_dyn()                                   # beam search forks a builder (backend.rs:205)
    g.fuse_until_closed()
    assert len(f) == 1 and len(g) == 3
    assert len(f.finish()) == 1 and len(g.finish()) == 3
    assert g.finish().name == "ElementWise"
    st.sync()
    assert v.shape == (8, 8)


def test_optimization_executes_on_the_queue_head_and_rejects_other_operations(st):
    a = st.placeholder((16, 16))
    t = a.exp()
    u = t.mul_scalar(3.0); t.drop()
    opt = st.fuser(F.BLOCK_ELEMWISE).fuse_until_closed().finish()
    opt.execute(st)                                     # pops the three entries it stands for
    (blk,) = st.blocks()
    assert (blk.kind, blk.n_ops, blk.from_cache) == (F.BLOCK_ELEMWISE, 2, 0)
    w = u.log()                                         # a different head: the optimization must refuse it
    with pytest.raises(abi.B200Error):
        opt.execute(st)
    st.sync()
    assert w.shape == (16, 16)


def test_optimization_state_round_trip_runs_at_other_sizes():
    s1, s2 = F.FusionStream(plan_only=True), F.FusionStream(plan_only=True)
    try:
        def record(s, n, m, k):
            x, w, bias = s.placeholder((n, k)), s.placeholder((k, m)), s.placeholder((1, m))
            h = x.matmul(w)
            hb = h.add(bias); h.drop()
            return F.gelu(hb), hb
        y1, hb1 = record(s1, 64, 128, 32)
        hb1.drop()
        opt = s1.fuser(F.BLOCK_MATMUL).fuse_until_closed().finish()
        state = opt.to_state()
        assert isinstance(state, bytes) and len(state) > 64
        again = F.Optimization.from_state(state)
        assert again.to_state() == state and len(again) == len(opt) == 7 and again.name == "Matmul"
        # the state holds relative ids only: it binds to the same graph at other concrete sizes, on another stream
        y2, hb2 = record(s2, 256, 512, 96)
        hb2.drop()
        again.execute(s2)
        (blk,) = s2.blocks()
        assert (blk.kind, blk.n_ops, blk.n_inputs) == (F.BLOCK_MATMUL, 7, 3)
        assert y2.shape == (256, 512)
        for bad in (b"", b"B2OP", state[:-3], state[:40] + bytes(len(state) - 40)):
            with pytest.raises(abi.B200Error):
                F.Optimization.from_state(bad)
    finally:
        s1.close(); s2.close()


# ---------------------------------------------------------------- relative form + plan cache
def _mlp_step(s, n, d, scale):
    x, w = s.placeholder((n, d)), s.placeholder((d, d))
    h = x.matmul(w)
    a = h.mul_scalar(scale); h.drop()
    y = F.softmax(a, 1); a.drop()
    r = y.sum_dim(0)
    return y, r


def test_plan_cache_reexecutes_with_new_shapes_and_scalars(st):
    y, r = _mlp_step(st, 32, 64, 0.5)
    st.sync()
    first = [(b.kind, b.n_ops, b.from_cache) for b in st.blocks()]
    assert [f[2] for f in first] == [0] * len(first)
    cs = st.cache_stats()
    assert (cs.hits, cs.misses, cs.plans) == (0, 1, 1)
    st.clear_blocks()
    y2, r2 = _mlp_step(st, 48, 128, 0.25)               # same graph, other extents and another scalar
    st.sync()
    again = [(b.kind, b.n_ops, b.from_cache) for b in st.blocks()]
    assert [(k, n) for k, n, _ in again] == [(k, n) for k, n, _ in first]
    assert all(c == 1 for _, _, c in again)             # served from the store: no fuser ran
    cs = st.cache_stats()
    assert (cs.hits, cs.misses, cs.plans) == (1, 1, 1)
    assert y2.shape == (48, 128) and r2.shape == (1, 128)


def test_relative_shapes_distinguish_broadcast_and_equal_extents(st):
    a, b = st.placeholder((16, 16)), st.placeholder((16, 16))
    a.add(b)
    st.sync()
    c, d = st.placeholder((16, 32)), st.placeholder((16, 32))     # 16 ≠ 32: another relative trace
    c.add(d)
    st.sync()
    e, f = st.placeholder((8, 8)), st.placeholder((8, 8))         # square again: the first plan
    e.add(f)
    st.sync()
    g, h = st.placeholder((8, 8)), st.placeholder((1, 8))         # broadcast row: extent 1 is always shape id 0
    g.add(h)
    st.sync()
    cs = st.cache_stats()
    assert (cs.hits, cs.misses, cs.plans) == (1, 3, 3)


def test_plan_cache_key_carries_layout_facts_the_fusers_used(st):
    # softmax along the last axis of a dense tensor is row-resident; of a transposed view it is not —
    # same ops, same extents, different plan
    x = st.placeholder((64, 64))
    F.softmax(x, 1)
    st.sync()
    assert [b.kind for b in st.blocks()] == [F.BLOCK_ROWNORM]
    st.clear_blocks()
    xt = x.swap_dims(0, 1)
    F.softmax(xt, 1)
    st.sync()
    assert F.BLOCK_ROWNORM not in [b.kind for b in st.blocks()]
    assert all(b.from_cache == 0 for b in st.blocks())


# ---------------------------------------------------------------- in-place outputs, views, indexed reads
def test_consumed_input_is_reused_in_place(st):
    x, other = st.placeholder((128, 128)), st.placeholder((128, 128))
    before = st.cache_stats().inplace_aliases
    y = x.add(other); x.drop()                           # x is ReadWrite here: y may overwrite it
    st.sync()
    (blk,) = st.blocks()
    assert blk.aliased == 1 and st.cache_stats().inplace_aliases == before + 1
    st.clear_blocks()
    z = y.exp()                                          # y stays alive → a fresh output
    st.sync()
    assert st.blocks()[0].aliased == 0
    st.clear_blocks()
    row = st.placeholder((1, 128))
    w = row.add(z); row.drop()                           # a broadcast input cannot hold the full-size output
    st.sync()
    assert st.blocks()[0].aliased == 0 and w.shape == (128, 128)
    st.clear_blocks()
    v = z.swap_dims(0, 1)
    u = z.mul_scalar(2.0); z.drop()                      # the buffer has a second owner (the view)
    st.sync()
    assert st.blocks()[0].aliased == 0 and v.shape == u.shape


def test_views_are_metadata_only(st):
    x = st.placeholder((4, 6, 8))
    r = x.reshape((24, 8))
    e = st.placeholder((1, 8)).expand((24, 8))
    sl = x.slice([(1, 3), (0, 6), (2, 6)])
    sw = x.swap_dims(0, 2)
    assert (r.shape, e.shape, sl.shape, sw.shape) == ((24, 8), (24, 8), (2, 6, 4), (8, 6, 4))
    st.sync()
    assert st.blocks() == []                             # views of materialised tensors never reach the queue
    y = r.add(e)
    st.sync()
    (blk,) = st.blocks()
    assert (blk.kind, blk.n_inputs, blk.launches) == (F.BLOCK_ELEMWISE, 2, 0)
    with pytest.raises(abi.B200Error):
        x.reshape((5, 5))
    with pytest.raises(abi.B200Error):
        x.slice([(0, 5), (0, 6), (0, 8)])
    with pytest.raises(abi.B200Error):
        x.expand((4, 6, 9))
    assert y.shape == (24, 8)


def test_lone_view_of_a_pending_tensor_is_its_own_block(st):
    # tests/fusion/fusion_shape.rs: a view is fused only as an INPUT of a later kernel, never on its own
    a, b = st.placeholder((8, 16)), st.placeholder((8, 16))
    t = a.add(b)
    v = t.reshape((16, 8))
    u = v.exp()
    st.sync()
    assert [(b.kind, b.launches) for b in st.blocks()] == [(F.BLOCK_ELEMWISE, 0), (F.BLOCK_VIEW, 0), (F.BLOCK_ELEMWISE, 0)]
    st.clear_blocks()
    w = u.swap_dims(0, 1).reshape((128,))                 # reshape of a strided view has to copy first
    st.sync()
    (blk,) = st.blocks()
    assert blk.kind == F.BLOCK_VIEW and blk.launches == 1
    assert w.shape == (128,)


def test_gather_and_select_run_as_their_own_block(st):
    x = st.placeholder((10, 32))
    idx = st.placeholder((4,), abi.I64)
    gi = st.placeholder((10, 5), abi.I64)
    s = x.select(0, idx)
    g = x.gather(1, gi)
    y = s.exp()
    st.sync()
    assert [(b.kind, b.n_ops) for b in st.blocks()] == [(F.BLOCK_EAGER, 1), (F.BLOCK_EAGER, 1), (F.BLOCK_ELEMWISE, 1)]
    assert s.shape == (4, 32) and g.shape == (10, 5) and y.shape == (4, 32)
    with pytest.raises(abi.B200Error):
        x.select(0, gi)
    with pytest.raises(abi.B200Error):
        x.gather(2, gi)


# ---------------------------------------------------------------- a whole module graph through the stream
def test_encoder_layer_op_stream_is_carved_into_the_expected_kernels(st):
    """configs[3]'s encoder forward as burn-nn's primitive op stream (burn_b200/stream_model.py): every Linear is one
    Matmul block with its bias (and gelu / scale / the residual add that follows) as the epilogue, softmax and both
    layer_norms are row-resident blocks, views launch nothing — and the second forward comes from the plan store."""
    from collections import Counter
    from burn_b200.stream_model import StreamEncoder
    enc = StreamEncoder.placeholders(st, 64, 256, 4, 2)
    x = st.placeholder((8, 32, 64))
    y = enc.forward(x)
    st.sync()
    blocks = st.blocks()
    per_kind = Counter(b.kind for b in blocks)
    assert per_kind[F.BLOCK_MATMUL] == 2 * 8              # q, k, v, scores, context, out, ff1, ff2 per layer
    assert per_kind[F.BLOCK_ROWNORM] == 2 * 3             # softmax, ln1, ln2
    assert per_kind[F.BLOCK_ELEMWISE] == 0                # the residual adds ride the out / ff2 GEMM epilogues
    assert per_kind[F.BLOCK_REDUCE] == 0 and per_kind[F.BLOCK_EAGER] == 0
    mm = [b for b in blocks if b.kind == F.BLOCK_MATMUL]
    # context·v alone; q, k, v + bias; scores + scale; out and ff2 + bias + residual; ff1 + bias + 5 gelu primitives
    assert sorted(b.n_ops for b in mm[:8]) == [1, 2, 2, 2, 2, 3, 3, 7]
    assert sum(b.n_ops for b in blocks) == 2 * (22 + 5 + 2 * 9 + 9)   # every recorded op landed in exactly one block
    assert sum(b.launches for b in blocks if b.kind == F.BLOCK_VIEW) == 2  # only the context reshape copies
    assert y.shape == (8, 32, 64)
    st.clear_blocks()
    y2 = enc.forward(st.placeholder((6, 24, 64)))         # other batch and sequence length: same relative graph
    st.sync()
    assert len(st.blocks()) == len(blocks) and all(b.from_cache for b in st.blocks())
    assert y2.shape == (6, 24, 64)
    st.clear_blocks()
    enc.forward(st.placeholder((4, 16, 64)))              # batch == heads and seq == head dim: extents coincide, so the
    st.sync()                                             # relative shapes differ — a new plan, exactly as in the reference
    assert not any(b.from_cache for b in st.blocks())
