"""GPU parity: gather / scatter_add / select / select_add / copy-cast / arange / random.

Protocol of crates/burn-backend-tests/tests/cubecl/{gather,scatter,select,select_assign}.rs
and the golden values of tests/tensor/float/ops/{gather_scatter,select}.rs.  Indexing results
are bit-exact; scatter_add / select_add keep the oracle's sequential accumulation order
(crates/burn-ndarray/src/ops/base.rs:140-183), so float sums are bit-exact too.
"""
import ctypes as C

import numpy as np
import pytest

from burn_b200 import _abi as abi
from burn_b200 import device as dv
from burn_b200.device import DeviceTensor
from oracle import oracle
from tests import helpers as H

pytestmark = pytest.mark.gpu


def lib():
    return abi.load()


def gather(dim, t, idx):
    out = DeviceTensor.empty(idx.shape, t.dtype)
    a, b, c = t.desc(), idx.desc(), out.desc()
    abi.check(lib().b200_launch_gather(dim, C.byref(a), C.byref(b), C.byref(c), None))
    return out.numpy()


def scatter_add(dim, t, idx, v):
    a, b, c = t.desc(), idx.desc(), v.desc()
    abi.check(lib().b200_launch_scatter_add(dim, C.byref(a), C.byref(b), C.byref(c), None))
    return t.numpy()


def select(dim, t, idx):
    shape = list(t.shape)
    shape[dim] = idx.shape[0]
    out = DeviceTensor.empty(shape, t.dtype)
    a, b, c = t.desc(), idx.desc(), out.desc()
    abi.check(lib().b200_launch_select(dim, C.byref(a), C.byref(b), C.byref(c), None))
    return out.numpy()


def select_add(dim, t, idx, v):
    a, b, c = t.desc(), idx.desc(), v.desc()
    abi.check(lib().b200_launch_select_add(dim, C.byref(a), C.byref(b), C.byref(c), None))
    return t.numpy()


def test_gather_reference_goldens(dev):
    # crates/burn-backend-tests/tests/tensor/float/ops/gather_scatter.rs (should_gather_1d_dim0 / 2d)
    t = H.up(np.array([0.0, 1.0, 2.0], dtype=np.float32))
    idx = H.up(np.array([1, 1, 0, 1, 2], dtype=np.int64))
    H.assert_exact(gather(0, t, idx), np.array([1.0, 1.0, 0.0, 1.0, 2.0], dtype=np.float32))
    t2 = H.up(np.array([[0.0, 1.0, 2.0], [3.0, 4.0, 5.0]], dtype=np.float32))
    i0 = H.up(np.array([[0, 1, 0], [1, 0, 1]], dtype=np.int64))
    H.assert_exact(gather(0, t2, i0), np.array([[0.0, 4.0, 2.0], [3.0, 1.0, 5.0]], dtype=np.float32))
    i1 = H.up(np.array([[2, 1, 0, 0], [2, 0, 1, 2]], dtype=np.int64))
    H.assert_exact(gather(1, t2, i1), np.array([[2.0, 1.0, 0.0, 0.0], [5.0, 3.0, 4.0, 5.0]], dtype=np.float32))


@pytest.mark.parametrize("shape,dim", [((7, 33), 1), ((7, 33), 0), ((4, 6, 20), 1), ((128, 1000), 1)])
@pytest.mark.parametrize("itype", [np.int32, np.int64])
def test_gather_random_vs_oracle(dev, shape, dim, itype):
    rng = np.random.default_rng(1)
    t = rng.uniform(-1, 1, shape).astype(np.float32)
    ishape = list(shape)
    ishape[dim] = 5
    idx = rng.integers(0, shape[dim], size=ishape).astype(itype)
    H.assert_exact(gather(dim, H.up(t), H.up(idx)), oracle.float_gather(dim, t, idx))


def test_cross_entropy_gather_shape(dev):
    # CE loss: log-probs [N, V] gathered by targets [N, 1] (crates/burn-nn/src/loss/cross_entropy.rs:171-197)
    rng = np.random.default_rng(2)
    lp = rng.uniform(-10, 0, (512, 5000)).astype(np.float32)
    tg = rng.integers(0, 5000, size=(512, 1)).astype(np.int64)
    H.assert_exact(gather(1, H.up(lp), H.up(tg)), oracle.float_gather(1, lp, tg))


def test_scatter_add_reference_goldens(dev):
    # gather_scatter.rs should_scatter_1d / 2d_dim0
    t = H.up(np.zeros(3, dtype=np.float32))
    H.assert_exact(scatter_add(0, t, H.up(np.array([1, 0, 2], dtype=np.int64)),
                               H.up(np.array([5.0, 4.0, 3.0], dtype=np.float32))),
                   np.array([4.0, 5.0, 3.0], dtype=np.float32))
    t = H.up(np.zeros((2, 3), dtype=np.float32))
    v = H.up(np.array([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]], dtype=np.float32))
    i = H.up(np.array([[1, 0, 1], [1, 1, 0]], dtype=np.int64))
    H.assert_exact(scatter_add(0, t, i, v), np.array([[0.0, 2.0, 6.0], [5.0, 5.0, 3.0]], dtype=np.float32))


@pytest.mark.parametrize("shape,dim,n", [((50, 8), 0, 300), ((6, 40), 1, 100), ((3, 64, 5), 1, 200)])
def test_scatter_add_bit_exact_with_collisions(dev, shape, dim, n):
    rng = np.random.default_rng(3)
    t = rng.uniform(-1, 1, shape).astype(np.float32)
    ishape = list(shape)
    ishape[dim] = n
    idx = rng.integers(0, shape[dim], size=ishape).astype(np.int64)
    v = rng.uniform(-1, 1, ishape).astype(np.float32)
    H.assert_exact(scatter_add(dim, H.up(t), H.up(idx), H.up(v)), oracle.float_scatter_add(dim, t, idx, v))


def test_select_and_embedding_forward(dev):
    # select.rs goldens + embedding forward = select(weight [V, d], ids) (ops/modules/base.rs:140-160)
    t = np.array([[0.0, 1.0, 2.0], [3.0, 4.0, 5.0]], dtype=np.float32)
    H.assert_exact(select(0, H.up(t), H.up(np.array([1, 0], dtype=np.int64))), t[[1, 0]])
    H.assert_exact(select(1, H.up(t), H.up(np.array([1, 1, 0, 1, 2], dtype=np.int64))), t[:, [1, 1, 0, 1, 2]])
    rng = np.random.default_rng(4)
    w = rng.standard_normal((1000, 64)).astype(np.float32)
    ids = rng.integers(0, 1000, size=(777,)).astype(np.int32)
    H.assert_exact(select(0, H.up(w), H.up(ids)), oracle.float_select(w, 0, ids))


def test_select_add_embedding_backward_bit_exact(dev):
    # embedding backward = zeros[V, d].select_add(0, ids, grad) (ops/modules/base.rs:161-180)
    rng = np.random.default_rng(5)
    V, d, n = 300, 48, 2000
    ids = rng.integers(0, V, size=(n,)).astype(np.int64)
    g = rng.standard_normal((n, d)).astype(np.float32)
    got = select_add(0, H.up(np.zeros((V, d), dtype=np.float32)), H.up(ids), H.up(g))
    H.assert_exact(got, oracle.float_select_add(np.zeros((V, d), dtype=np.float32), 0, ids, g))
    # select.rs should_select_add_2d_dim1-style case
    t = np.array([[0.0, 1.0, 2.0], [3.0, 4.0, 5.0]], dtype=np.float32)
    v = np.array([[1.0, 2.0, 3.0, 4.0, 5.0], [6.0, 7.0, 8.0, 9.0, 10.0]], dtype=np.float32)
    i = np.array([1, 1, 0, 1, 2], dtype=np.int64)
    H.assert_exact(select_add(1, H.up(t), H.up(i), H.up(v)), oracle.float_select_add(t, 1, i, v))


def test_copy_of_strided_views_and_casts(dev):
    rng = np.random.default_rng(6)
    x = rng.uniform(-100, 100, (6, 10, 12)).astype(np.float32)
    t = H.up(x)
    H.assert_exact(t.permute([2, 0, 1]).contiguous().numpy(), x.transpose(2, 0, 1))
    H.assert_exact(t.swap_dims(0, 2).contiguous().numpy(), x.swapaxes(0, 2))
    H.assert_exact(t.slice([(1, 5), (2, 9), (3, 11)]).contiguous().numpy(), x[1:5, 2:9, 3:11])
    H.assert_exact(H.up(x[:1]).expand((4, 10, 12)).contiguous().numpy(), np.broadcast_to(x[:1], (4, 10, 12)))
    out = DeviceTensor.empty(x.shape, abi.I32)
    s, d = t.desc(), out.desc()
    abi.check(lib().b200_launch_copy(C.byref(s), C.byref(d), None))
    H.assert_exact(out.numpy(), x.astype(np.int32))
    outh = DeviceTensor.empty(x.shape, abi.F16)
    d = outh.desc()
    abi.check(lib().b200_launch_copy(C.byref(s), C.byref(d), None))
    H.assert_exact(outh.numpy(), x.astype(np.float16))


def test_arange(dev):
    out = DeviceTensor.empty((1000,), abi.I64)
    d = out.desc()
    abi.check(lib().b200_launch_arange(C.byref(d), 5, 3, None))
    H.assert_exact(out.numpy(), np.arange(1000, dtype=np.int64) * 3 + 5)


def test_random_distributions(dev):
    # statistical checks only, like tests/cubecl/{uniform,normal,bernoulli}.rs
    n = 1 << 20
    out = DeviceTensor.empty((n,))
    d = out.desc()
    abi.check(lib().b200_launch_random(C.byref(d), 0, -2.0, 3.0, 42, 0, None))
    u = out.numpy()
    assert u.min() >= -2.0 and u.max() < 3.0
    assert abs(u.mean() - 0.5) < 0.01 and abs(u.var() - 25 / 12) < 0.02
    abi.check(lib().b200_launch_random(C.byref(d), 1, 1.0, 2.0, 43, 0, None))
    g = out.numpy()
    assert abs(g.mean() - 1.0) < 0.01 and abs(g.std() - 2.0) < 0.01
    abi.check(lib().b200_launch_random(C.byref(d), 2, 0.25, 0.0, 44, 0, None))
    b = out.numpy()
    assert set(np.unique(b)) <= {0.0, 1.0} and abs(b.mean() - 0.25) < 0.005
    # same seed → same stream; different offset → different values
    abi.check(lib().b200_launch_random(C.byref(d), 0, 0.0, 1.0, 7, 0, None))
    r1 = out.numpy()
    abi.check(lib().b200_launch_random(C.byref(d), 0, 0.0, 1.0, 7, 0, None))
    assert np.array_equal(r1, out.numpy())
    abi.check(lib().b200_launch_random(C.byref(d), 0, 0.0, 1.0, 7, n, None))
    assert not np.array_equal(r1, out.numpy())


def test_gather_shape_mismatch_is_an_error(dev):
    t = H.up(np.zeros((2, 3), dtype=np.float32))
    idx = H.up(np.zeros((3, 3), dtype=np.int64))
    out = DeviceTensor.empty((3, 3))
    a, b, c = t.desc(), idx.desc(), out.desc()
    assert lib().b200_launch_gather(1, C.byref(a), C.byref(b), C.byref(c), None) == abi.ERR_SHAPE


@pytest.mark.parametrize("what", ["gather", "select", "scatter_add", "select_add", "select_add_rows"])
def test_out_of_range_indices_fail_loudly(dev, what):
    """The reference panics on an index >= the axis length (crates/burn-ndarray/src/ops/base.rs:106-183 index
    with `as usize`).  A kernel cannot panic: the access is skipped (no out-of-bounds read or write — rows = 10
    makes the chunked scatter kernels' rounded-up range [0, 12) cover the bad index) and the next synchronising
    call reports B200_ERR_SHAPE."""
    rows = 10
    t = H.up(np.arange(rows * 8, dtype=np.float32).reshape(rows, 8))
    before = t.numpy().copy()
    if what == "gather":
        idx = H.up(np.full((3, 8), 11, dtype=np.int64))
        out = DeviceTensor.empty((3, 8))
        a, b, c = t.desc(), idx.desc(), out.desc()
        abi.check(lib().b200_launch_gather(0, C.byref(a), C.byref(b), C.byref(c), None))
    elif what == "select":
        idx = H.up(np.array([0, 10, -1], dtype=np.int32))
        out = DeviceTensor.empty((3, 8))
        a, b, c = t.desc(), idx.desc(), out.desc()
        abi.check(lib().b200_launch_select(0, C.byref(a), C.byref(b), C.byref(c), None))
    elif what == "scatter_add":
        idx = H.up(np.full((2, 8), 11, dtype=np.int64))
        v = H.up(np.ones((2, 8), dtype=np.float32))
        a, b, c = t.desc(), idx.desc(), v.desc()
        abi.check(lib().b200_launch_scatter_add(0, C.byref(a), C.byref(b), C.byref(c), None))
    else:
        tt = t if what == "select_add_rows" else H.up(before.T.copy()).swap_dims(0, 1)   # strided: generic kernel
        idx = H.up(np.array([11, 3], dtype=np.int64))
        v = H.up(np.ones((2, 8), dtype=np.float32))
        a, b, c = tt.desc(), idx.desc(), v.desc()
        abi.check(lib().b200_launch_select_add(0, C.byref(a), C.byref(b), C.byref(c), None))
    with pytest.raises(abi.B200Error) as e:
        dv.sync()
    assert e.value.status == abi.ERR_SHAPE and "out of range" in e.value.message
    dv.sync()                                   # the flag is sticky until reported, then cleared
    if what == "scatter_add":
        H.assert_exact(t.numpy(), before)       # nothing was written past (or into) the table
    if what == "select_add_rows":
        want = before.copy()
        want[3] += 1.0
        H.assert_exact(t.numpy(), want)         # the valid index still landed


def test_retain_free_refcount(dev):
    """b200_alloc / b200_retain / b200_free give Handle::can_mut's refcount semantics
    (crates/burn-ir/src/handle.rs:92-111): a clone is a refcount bump, memory goes back at zero."""
    p = C.c_void_p()
    abi.check(lib().b200_alloc(C.byref(p), 4096, None))
    n = C.c_uint32()
    abi.check(lib().b200_refcount(p, C.byref(n)))
    assert n.value == 1                         # sole owner: may be mutated in place
    abi.check(lib().b200_retain(p))
    abi.check(lib().b200_refcount(p, C.byref(n)))
    assert n.value == 2
    abi.check(lib().b200_free(p, None))         # drops one owner, memory stays
    abi.check(lib().b200_refcount(p, C.byref(n)))
    assert n.value == 1
    abi.check(lib().b200_memset(p, 0, 4096, None))
    abi.check(lib().b200_free(p, None))
    assert lib().b200_refcount(p, C.byref(n)) == abi.ERR_INVALID
    dv.sync()


# ---- data movement: cat / slice / slice_assign / flip / repeat_dim
# (crates/burn-backend-tests/tests/tensor/float/ops/{cat,slice,slice_assign,flip,repeat_dim}.rs; copies → bit-exact)
def test_cat_matches_oracle_on_strided_and_empty_inputs(dev):
    from burn_b200 import ops
    rng = np.random.default_rng(11)
    a = rng.uniform(-1, 1, (37, 5, 130)).astype(np.float32)
    b = rng.uniform(-1, 1, (37, 9, 130)).astype(np.float32)
    c = rng.uniform(-1, 1, (130, 3, 37)).astype(np.float32)          # joins as a permuted (strided) view
    e = np.zeros((37, 0, 130), dtype=np.float32)
    got = ops.float_cat([H.up(a), H.up(e), H.up(b), H.up(c).permute([2, 1, 0])], 1).numpy()
    H.assert_exact(got, oracle.float_cat([a, e, b, np.transpose(c, (2, 1, 0))], 1))
    for dim in (0, 2, -1):
        H.assert_exact(ops.float_cat([H.up(a), H.up(a)], dim).numpy(), oracle.float_cat([a, a], dim % 3))
    with pytest.raises(abi.B200Error):
        ops.float_cat([H.up(a), H.up(c)], 1)                          # cat.rs:47-56: mismatched dims must fail


def test_slice_and_slice_assign_match_oracle(dev):
    from burn_b200 import ops
    rng = np.random.default_rng(12)
    x = rng.uniform(-1, 1, (64, 33, 70)).astype(np.float32)
    v = rng.uniform(-1, 1, (10, 33, 21)).astype(np.float32)
    ranges = [(5, 15), (0, 33), (-30, -9)]
    H.assert_exact(ops.float_slice(H.up(x), ranges).contiguous().numpy(), oracle.float_slice(x, ranges))
    dx = H.up(x)
    got = ops.float_slice_assign(dx, ranges, H.up(v)).numpy()
    H.assert_exact(got, oracle.float_slice_assign(x, ranges, v))
    H.assert_exact(dx.numpy(), x)                                     # the still-referenced input is untouched
    # clamp_when_slice_exceeds_dimension (slice.rs:362) and a reduction reading the view in place
    H.assert_exact(ops.float_slice(H.up(x), [(60, 999)]).contiguous().numpy(), x[60:])
    s = ops.float_sum_dim(ops.float_slice(H.up(x), [(0, 64), (3, 20)]), 2).numpy()
    assert oracle.approx_eq_mask(s, oracle.float_sum_dim(np.ascontiguousarray(x[:, 3:20]), 2), H.REL_REDUCE, 1e-6).all()
    with pytest.raises(abi.B200Error):
        ops.float_slice_assign(H.up(x), ranges, H.up(v[:, :, :20]))


@pytest.mark.parametrize("axes", [[0], [2], [0, 1, 2], [1, -1]])
def test_flip_matches_oracle(dev, axes):
    from burn_b200 import ops
    x = np.random.default_rng(13).uniform(-1, 1, (19, 8, 257)).astype(np.float32)
    H.assert_exact(ops.float_flip(H.up(x), axes).numpy(), oracle.float_flip(x, [a % 3 for a in axes]))
    xt = H.up(x).swap_dims(0, 2)
    H.assert_exact(ops.float_flip(xt, [0]).numpy(), oracle.float_flip(np.swapaxes(x, 0, 2), [0]))


@pytest.mark.parametrize("dim,times", [(0, 3), (1, 1), (2, 4)])
def test_repeat_dim_matches_oracle(dev, dim, times):
    from burn_b200 import ops
    x = np.random.default_rng(14).uniform(-1, 1, (6, 1 if dim == 1 else 7, 65)).astype(np.float32)
    H.assert_exact(ops.float_repeat_dim(H.up(x), dim, times).numpy(), oracle.float_repeat_dim(x, dim, times))
