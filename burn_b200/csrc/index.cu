// Data movement and indexing kernels: strided copy / cast, gather, scatter_add,
// select, select_add, arange and Philox random.
//
// Reference kernels replaced: crates/burn-cubecl/src/kernel/contiguous.rs,
// kernel/cast/base.rs:13, kernel/index/{gather.rs:18, scatter.rs:14, select.rs:12,
// select_assign.rs:14}, ops/numeric.rs (arange/full), kernel/prng/*.
// Oracle semantics: crates/burn-ndarray/src/ops/base.rs:106-183 (gather / scatter
// are sequential along `dim`; we keep that order so float accumulation is
// bit-identical).
#include "tape_host.cuh"

namespace b200 {

// Out-of-range indices: the reference panics (crates/burn-ndarray/src/ops/base.rs:106-183 index with
// `as usize`); here the offending access is skipped, a sticky code lands in a host-mapped flag and the
// next synchronising call (b200_stream_sync / b200_device_sync / blocking d2h copy) returns B200_ERR_SHAPE.
struct IdxTensor {
  void *ptr;
  int64_t shape[kMaxDims];
  int64_t strides[kMaxDims];
  int32_t dtype;
  int32_t rank;
};

static IdxTensor to_idx(const b200_tensor &t) {
  IdxTensor r;
  r.ptr = t.ptr;
  r.dtype = t.dtype;
  r.rank = t.rank;
  for (int d = 0; d < kMaxDims; ++d) {
    r.shape[d] = d < t.rank ? t.shape[d] : 1;
    r.strides[d] = d < t.rank ? t.strides[d] : 0;
  }
  return r;
}

__device__ __forceinline__ int64_t load_index(const void *p, int32_t dtype, int64_t off) {
  return dtype == B200_I64 ? reinterpret_cast<const int64_t *>(p)[off]
                           : (int64_t) reinterpret_cast<const int32_t *>(p)[off];
}

__device__ __forceinline__ void copy_elem(void *dst, int64_t doff, const void *src, int64_t soff, int es) {
  switch (es) {
    case 1: reinterpret_cast<uint8_t *>(dst)[doff] = reinterpret_cast<const uint8_t *>(src)[soff]; break;
    case 2: reinterpret_cast<uint16_t *>(dst)[doff] = reinterpret_cast<const uint16_t *>(src)[soff]; break;
    case 4: reinterpret_cast<uint32_t *>(dst)[doff] = reinterpret_cast<const uint32_t *>(src)[soff]; break;
    default: reinterpret_cast<uint64_t *>(dst)[doff] = reinterpret_cast<const uint64_t *>(src)[soff]; break;
  }
}

// out[c] = in[c with c[dim] = idx[c]]   (one thread per output element)
__global__ void gather_kernel(IdxTensor in, IdxTensor idx, IdxTensor out, int dim, int es, int64_t n, int32_t *err) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t rest = i, ioff = 0, xoff = 0, ooff = 0;
    int64_t cdim = 0;
#pragma unroll
    for (int d = kMaxDims - 1; d >= 0; --d) {
      if (d < out.rank) {
        const int64_t c = rest % out.shape[d];
        rest /= out.shape[d];
        xoff += c * idx.strides[d];
        ooff += c * out.strides[d];
        if (d == dim) cdim = c; else ioff += c * in.strides[d];
      }
    }
    (void)cdim;
    const int64_t k = load_index(idx.ptr, idx.dtype, xoff);
    if (k < 0 || k >= in.shape[dim]) { *err = kIdxErrGather; continue; }   // the reference panics; reported at the next sync
    copy_elem(out.ptr, ooff, in.ptr, ioff + k * in.strides[dim], es);
  }
}

// out[o, i, c] = in[o, idx[i], c]  (select / index_select; idx is 1-D)
__global__ void select_kernel(IdxTensor in, IdxTensor idx, IdxTensor out, int dim, int es, int64_t n, int32_t *err) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t rest = i, ioff = 0, ooff = 0, k = 0;
#pragma unroll
    for (int d = kMaxDims - 1; d >= 0; --d) {
      if (d < out.rank) {
        const int64_t c = rest % out.shape[d];
        rest /= out.shape[d];
        ooff += c * out.strides[d];
        if (d == dim) k = load_index(idx.ptr, idx.dtype, c * idx.strides[0]);
        else ioff += c * in.strides[d];
      }
    }
    if (k < 0 || k >= in.shape[dim]) { *err = kIdxErrSelect; continue; }
    copy_elem(out.ptr, ooff, in.ptr, ioff + k * in.strides[dim], es);
  }
}

__device__ __forceinline__ void add_elem(void *t, int64_t off, int32_t tdt, const void *v, int64_t voff,
                                         int32_t vdt) {
  if (tdt == B200_F32) {
    float *p = reinterpret_cast<float *>(t) + off;
    *p = __fadd_rn(*p, reinterpret_cast<const float *>(v)[voff]);
  } else if (tdt == B200_I32) {
    reinterpret_cast<int32_t *>(t)[off] += reinterpret_cast<const int32_t *>(v)[voff];
  } else if (tdt == B200_I64) {
    reinterpret_cast<int64_t *>(t)[off] += reinterpret_cast<const int64_t *>(v)[voff];
  } else if (tdt == B200_BOOL || tdt == B200_U8) {  // scatter_or
    reinterpret_cast<uint8_t *>(t)[off] |= reinterpret_cast<const uint8_t *>(v)[voff];
  } else if (tdt == B200_BF16) {
    __nv_bfloat16 *p = reinterpret_cast<__nv_bfloat16 *>(t) + off;
    *p = __float2bfloat16_rn(__bfloat162float(*p) + __bfloat162float(reinterpret_cast<const __nv_bfloat16 *>(v)[voff]));
  } else {
    __half *p = reinterpret_cast<__half *>(t) + off;
    *p = __float2half_rn(__half2float(*p) + __half2float(reinterpret_cast<const __half *>(v)[voff]));
  }
  (void)vdt;
}

// scatter_add: one thread per lane (all coords except `dim`), sequential along dim.
// The destination rows are partitioned into `chunks` ranges so that several
// threads can share a lane without ever touching the same address.
__global__ void scatter_add_kernel(IdxTensor t, IdxTensor idx, IdxTensor v, int dim, int64_t lanes,
                                   int chunks, int64_t rows_per_chunk, int32_t *err) {
  const int64_t total = lanes * chunks;
  for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < total;
       w += (int64_t)gridDim.x * blockDim.x) {
    const int64_t lane = w % lanes;
    const int chunk = (int)(w / lanes);
    const int64_t rows = t.shape[dim];
    const int64_t lo = chunk * rows_per_chunk, hi = min(rows, lo + rows_per_chunk);
    int64_t rest = lane, toff = 0, xoff = 0, voff = 0;
#pragma unroll
    for (int d = kMaxDims - 1; d >= 0; --d) {
      if (d < idx.rank && d != dim) {
        const int64_t c = rest % idx.shape[d];
        rest /= idx.shape[d];
        toff += c * t.strides[d];
        xoff += c * idx.strides[d];
        voff += c * v.strides[d];
      }
    }
    const int64_t n = idx.shape[dim];
    for (int64_t i = 0; i < n; ++i) {
      const int64_t k = load_index(idx.ptr, idx.dtype, xoff + i * idx.strides[dim]);
      if (chunk == 0 && (k < 0 || k >= rows)) *err = kIdxErrScatter;
      if (k >= lo && k < hi)
        add_elem(t.ptr, toff + k * t.strides[dim], t.dtype, v.ptr, voff + i * v.strides[dim], v.dtype);
    }
  }
}

// select_add: t[o, idx[i], c] += v[o, i, c], sequential in i per (o, c) lane.
__global__ void select_add_kernel(IdxTensor t, IdxTensor idx, IdxTensor v, int dim, int64_t lanes,
                                  int chunks, int64_t rows_per_chunk, int32_t *err) {
  const int64_t total = lanes * chunks;
  for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < total;
       w += (int64_t)gridDim.x * blockDim.x) {
    const int64_t lane = w % lanes;
    const int chunk = (int)(w / lanes);
    const int64_t rows = t.shape[dim];
    const int64_t lo = chunk * rows_per_chunk, hi = min(rows, lo + rows_per_chunk);
    int64_t rest = lane, toff = 0, voff = 0;
#pragma unroll
    for (int d = kMaxDims - 1; d >= 0; --d) {
      if (d < v.rank && d != dim) {
        const int64_t c = rest % v.shape[d];
        rest /= v.shape[d];
        toff += c * t.strides[d];
        voff += c * v.strides[d];
      }
    }
    const int64_t n = v.shape[dim];
    for (int64_t i = 0; i < n; ++i) {
      const int64_t k = load_index(idx.ptr, idx.dtype, i * idx.strides[0]);
      if (chunk == 0 && (k < 0 || k >= rows)) *err = kIdxErrSelectAdd;
      if (k >= lo && k < hi)
        add_elem(t.ptr, toff + k * t.strides[dim], t.dtype, v.ptr, voff + i * v.strides[dim], v.dtype);
    }
  }
}

// select_add on rows of a 2-D f32 table (embedding backward): t[idx[i], :] += v[i, :].
// A warp owns a chunk of target rows x 128 columns.  It scans the indices 32 at a time (one
// coalesced load), ballots the ones that land in its chunk and applies them in increasing i —
// the same per-element order as the sequential reference loop (crates/burn-ndarray/src/ops/base.rs
// select_assign), so the sums are bit-identical and need no atomics — while every index is read
// once per warp instead of once per thread.
__global__ void __launch_bounds__(256) select_add_rows_kernel(float *t, int64_t t_stride, const void *idx, int32_t idx_dtype,
                                                              int64_t idx_stride, const float *v, int64_t v_stride,
                                                              int64_t n, int64_t rows_t, int cols4, int col_groups,
                                                              int chunks, int64_t rows_per_chunk, int32_t *err) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t w = warp; w < (int64_t)col_groups * chunks; w += n_warps) {
    const int cg = (int)(w % col_groups);
    const int64_t lo = (w / col_groups) * rows_per_chunk, hi = min(rows_t, lo + rows_per_chunk);
    const int c4 = cg * 32 + lane;
    const bool col_ok = c4 < cols4;
    for (int64_t base = 0; base < n; base += 32) {
      const int64_t i = base + lane;
      const int64_t k = i < n ? load_index(idx, idx_dtype, i * idx_stride) : -1;
      if (w == 0 && i < n && (k < 0 || k >= rows_t)) *err = kIdxErrSelectAdd;
      unsigned hits = __ballot_sync(0xffffffffu, k >= lo && k < hi);
      while (hits) {
        const int b = __ffs(hits) - 1;
        hits &= hits - 1;
        const int64_t kb = __shfl_sync(0xffffffffu, k, b);
        if (col_ok) {
          float4 *dst = reinterpret_cast<float4 *>(t + kb * t_stride) + c4;
          const float4 a = *dst, x = __ldg(reinterpret_cast<const float4 *>(v + (base + b) * v_stride) + c4);
          *dst = make_float4(__fadd_rn(a.x, x.x), __fadd_rn(a.y, x.y), __fadd_rn(a.z, x.z), __fadd_rn(a.w, x.w));
        }
      }
    }
  }
}

// out[c] = in[c with c[d] -> shape[d] - 1 - c[d] on the flipped axes]  (bit mask `axes`)
__global__ void flip_kernel(IdxTensor in, IdxTensor out, uint32_t axes, int es, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t rest = i, ioff = 0, ooff = 0;
#pragma unroll
    for (int d = kMaxDims - 1; d >= 0; --d) {
      if (d < out.rank) {
        const int64_t c = rest % out.shape[d];
        rest /= out.shape[d];
        ooff += c * out.strides[d];
        ioff += (((axes >> d) & 1u) ? out.shape[d] - 1 - c : c) * in.strides[d];
      }
    }
    copy_elem(out.ptr, ooff, in.ptr, ioff, es);
  }
}

__global__ void arange_kernel(void *out, int32_t dtype, int64_t n, int64_t start, int64_t step) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t v = start + i * step;
    if (dtype == B200_I64) reinterpret_cast<int64_t *>(out)[i] = v;
    else if (dtype == B200_I32) reinterpret_cast<int32_t *>(out)[i] = (int32_t)v;
    else reinterpret_cast<float *>(out)[i] = (float)v;
  }
}

// ---- Philox4x32-10 ---------------------------------------------------------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  k[0] += 0x9E3779B9u;
  k[1] += 0xBB67AE85u;
}

__device__ __forceinline__ void philox4x32_10(uint64_t ctr, uint64_t seed, uint32_t (&out)[4]) {
  uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
  uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
  for (int r = 0; r < 10; ++r) philox_round(c, k);
#pragma unroll
  for (int j = 0; j < 4; ++j) out[j] = c[j];
}

__device__ __forceinline__ float u01(uint32_t x) {  // [0, 1)
  return (float)(x >> 8) * (1.0f / 16777216.0f);
}

__global__ void random_kernel(void *out, int32_t dtype, int64_t n, int kind, float lo, float hi,
                              uint64_t seed, uint64_t offset) {
  const int64_t n4 = (n + 3) / 4;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n4;
       q += (int64_t)gridDim.x * blockDim.x) {
    uint32_t r[4];
    philox4x32_10((uint64_t)q + offset, seed, r);
    float v[4];
    if (kind == 1) {  // normal via Box–Muller on two pairs
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const float u1 = 1.0f - u01(r[2 * p]);  // (0, 1]
        const float u2 = u01(r[2 * p + 1]);
        const float rad = sqrtf(-2.0f * logf(u1));
        float s, c;
        sincospif(2.0f * u2, &s, &c);
        v[2 * p] = lo + hi * rad * c;
        v[2 * p + 1] = lo + hi * rad * s;
      }
    } else if (kind == 2) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = u01(r[j]) < lo ? 1.0f : 0.0f;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = lo + (hi - lo) * u01(r[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t i = q * 4 + j;
      if (i >= n) break;
      uint32_t bits = (dtype == B200_F32 || dtype == B200_F16 || dtype == B200_BF16)
                          ? u_of(v[j])
                          : (uint32_t)(int32_t)v[j];
      store_one(out, dtype, i, bits);
    }
  }
}

static unsigned grid_for(int64_t n, int block) {
  const int64_t need = (n + block - 1) / block;
  const int64_t cap = (int64_t)sm_count() * 16;
  return (unsigned)std::max<int64_t>(1, std::min(need, cap));
}

static int32_t check_index_dtype(const b200_tensor *idx) {
  B200_REQUIRE(idx->dtype == B200_I32 || idx->dtype == B200_I64, B200_ERR_INVALID,
               "indices must be i32 or i64 (got dtype %d)", idx->dtype);
  return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" int32_t b200_launch_copy(const b200_tensor *src, const b200_tensor *dst, b200_stream s) {
  B200_REQUIRE(src && dst, B200_ERR_INVALID, "null argument");
  B200_REQUIRE(src->rank == dst->rank, B200_ERR_SHAPE, "copy rank mismatch %d vs %d", src->rank, dst->rank);
  auto cls = [](int32_t dt) { return (dt == B200_F32 || dt == B200_F16 || dt == B200_BF16) ? 0 : (dt == B200_BOOL ? 2 : 1); };
  const int cs = cls(src->dtype), cd = cls(dst->dtype);
  int op = B200_OP_MOV;
  if (cs == 0 && cd == 1) op = B200_OP_F2I;
  else if (cs == 1 && cd == 0) op = B200_OP_I2F;
  else if (cs == 2 && cd == 0) op = B200_OP_B2F;
  else if (cs == 2 && cd == 1) op = B200_OP_B2I;
  else if (cs == 0 && cd == 2) op = B200_OP_F2B;
  else if (cs == 1 && cd == 2) op = B200_OP_I2B;
  b200_tape_op t = {(uint8_t)op, (uint8_t)B200_ARG_INPUT(0), 0, 0, B200_DST_NONE, 0, {0, 0}};
  b200_tape tape = {&t, 1, nullptr, 0};
  return b200_launch_elemwise(&tape, src, 1, dst, 1, dst->rank, dst->shape, s);
}

extern "C" int32_t b200_launch_gather(int32_t dim, const b200_tensor *input, const b200_tensor *indices,
                                      const b200_tensor *out, b200_stream s) {
  B200_REQUIRE(input && indices && out, B200_ERR_INVALID, "null argument");
  B200_REQUIRE(input->rank == indices->rank && out->rank == indices->rank, B200_ERR_SHAPE, "gather rank mismatch");
  B200_REQUIRE(dim >= 0 && dim < input->rank, B200_ERR_SHAPE, "gather dim %d out of range", dim);
  B200_REQUIRE(input->dtype == out->dtype, B200_ERR_INVALID, "gather dtype mismatch");
  int32_t st = check_index_dtype(indices);
  if (st != B200_OK) return st;
  for (int d = 0; d < input->rank; ++d) {
    B200_REQUIRE(out->shape[d] == indices->shape[d], B200_ERR_SHAPE, "gather: out/indices shape mismatch at dim %d", d);
    if (d != dim)
      B200_REQUIRE(indices->shape[d] == input->shape[d], B200_ERR_SHAPE,
                   "gather: indices dim %d is %lld but tensor has %lld", d, (long long)indices->shape[d],
                   (long long)input->shape[d]);
  }
  const int64_t n = numel_of(out->shape, out->rank);
  if (n == 0) return B200_OK;
  gather_kernel<<<grid_for(n, 256), 256, 0, resolve_stream(s)>>>(to_idx(*input), to_idx(*indices), to_idx(*out),
                                                                 dim, dtype_size(input->dtype), n, index_error_flag());
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int32_t b200_launch_select(int32_t dim, const b200_tensor *input, const b200_tensor *indices,
                                      const b200_tensor *out, b200_stream s) {
  B200_REQUIRE(input && indices && out, B200_ERR_INVALID, "null argument");
  B200_REQUIRE(indices->rank == 1, B200_ERR_SHAPE, "select indices must be 1-D");
  B200_REQUIRE(input->rank == out->rank, B200_ERR_SHAPE, "select rank mismatch");
  B200_REQUIRE(dim >= 0 && dim < input->rank, B200_ERR_SHAPE, "select dim %d out of range", dim);
  B200_REQUIRE(input->dtype == out->dtype, B200_ERR_INVALID, "select dtype mismatch");
  int32_t st = check_index_dtype(indices);
  if (st != B200_OK) return st;
  for (int d = 0; d < input->rank; ++d)
    B200_REQUIRE(out->shape[d] == (d == dim ? indices->shape[0] : input->shape[d]), B200_ERR_SHAPE,
                 "select: bad output dim %d", d);
  const int64_t n = numel_of(out->shape, out->rank);
  if (n == 0) return B200_OK;
  select_kernel<<<grid_for(n, 256), 256, 0, resolve_stream(s)>>>(to_idx(*input), to_idx(*indices), to_idx(*out),
                                                                 dim, dtype_size(input->dtype), n, index_error_flag());
  B200_LAUNCH_CHECK();
  return B200_OK;
}

static void pick_chunks(int64_t lanes, int64_t rows, int &chunks, int64_t &rows_per_chunk) {
  const int64_t target = (int64_t)sm_count() * 2048;
  chunks = (int)std::max<int64_t>(1, std::min<int64_t>(rows, target / std::max<int64_t>(lanes, 1)));
  chunks = std::min(chunks, 1024);
  rows_per_chunk = (rows + chunks - 1) / chunks;
  chunks = (int)((rows + rows_per_chunk - 1) / rows_per_chunk);
}

extern "C" int32_t b200_launch_scatter_add(int32_t dim, const b200_tensor *tensor, const b200_tensor *indices,
                                           const b200_tensor *value, b200_stream s) {
  B200_REQUIRE(tensor && indices && value, B200_ERR_INVALID, "null argument");
  B200_REQUIRE(tensor->rank == indices->rank && value->rank == indices->rank, B200_ERR_SHAPE, "scatter rank mismatch");
  B200_REQUIRE(dim >= 0 && dim < tensor->rank, B200_ERR_SHAPE, "scatter dim %d out of range", dim);
  B200_REQUIRE(tensor->dtype == value->dtype, B200_ERR_INVALID, "scatter dtype mismatch");
  int32_t st = check_index_dtype(indices);
  if (st != B200_OK) return st;
  int64_t lanes = 1;
  for (int d = 0; d < tensor->rank; ++d) {
    B200_REQUIRE(indices->shape[d] == value->shape[d], B200_ERR_SHAPE, "scatter: indices/value shape mismatch at dim %d", d);
    if (d != dim) {
      B200_REQUIRE(indices->shape[d] <= tensor->shape[d], B200_ERR_SHAPE, "scatter: indices dim %d too large", d);
      lanes *= indices->shape[d];
    }
  }
  if (lanes == 0 || indices->shape[dim] == 0) return B200_OK;
  int chunks;
  int64_t rpc;
  pick_chunks(lanes, tensor->shape[dim], chunks, rpc);
  scatter_add_kernel<<<grid_for(lanes * chunks, 128), 128, 0, resolve_stream(s)>>>(
      to_idx(*tensor), to_idx(*indices), to_idx(*value), dim, lanes, chunks, rpc, index_error_flag());
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int32_t b200_launch_select_add(int32_t dim, const b200_tensor *tensor, const b200_tensor *indices,
                                          const b200_tensor *value, b200_stream s) {
  B200_REQUIRE(tensor && indices && value, B200_ERR_INVALID, "null argument");
  B200_REQUIRE(indices->rank == 1, B200_ERR_SHAPE, "select_add indices must be 1-D");
  B200_REQUIRE(tensor->rank == value->rank, B200_ERR_SHAPE, "select_add rank mismatch");
  B200_REQUIRE(dim >= 0 && dim < tensor->rank, B200_ERR_SHAPE, "select_add dim %d out of range", dim);
  B200_REQUIRE(tensor->dtype == value->dtype, B200_ERR_INVALID, "select_add dtype mismatch");
  int32_t st = check_index_dtype(indices);
  if (st != B200_OK) return st;
  int64_t lanes = 1;
  for (int d = 0; d < tensor->rank; ++d) {
    if (d == dim) {
      B200_REQUIRE(value->shape[d] == indices->shape[0], B200_ERR_SHAPE, "select_add: value dim %d != number of indices", d);
    } else {
      B200_REQUIRE(value->shape[d] == tensor->shape[d], B200_ERR_SHAPE, "select_add: value dim %d mismatch", d);
      lanes *= value->shape[d];
    }
  }
  if (lanes == 0 || indices->shape[0] == 0) return B200_OK;
  // embedding-backward shape: rows of a 2-D f32 table with aligned, unit-stride columns
  if (tensor->rank == 2 && dim == 0 && tensor->dtype == B200_F32 && tensor->shape[1] % 4 == 0 &&
      tensor->strides[1] == 1 && value->strides[1] == 1 && tensor->strides[0] % 4 == 0 && value->strides[0] % 4 == 0 &&
      ((uintptr_t)tensor->ptr % 16) == 0 && ((uintptr_t)value->ptr % 16) == 0 && tensor->shape[0] > 0) {
    const int cols4 = (int)(tensor->shape[1] / 4), col_groups = (cols4 + 31) / 32;
    const int64_t rows_t = tensor->shape[0];
    const int64_t want_warps = (int64_t)sm_count() * 16;
    const int chunks_r = (int)std::max<int64_t>(1, std::min<int64_t>(rows_t, want_warps / col_groups));
    const int64_t rpc_r = (rows_t + chunks_r - 1) / chunks_r;
    const int chunks_eff = (int)((rows_t + rpc_r - 1) / rpc_r);
    const int64_t warps = (int64_t)col_groups * chunks_eff;
    const unsigned grid = (unsigned)std::max<int64_t>(1, (warps + 7) / 8);
    select_add_rows_kernel<<<grid, 256, 0, resolve_stream(s)>>>(
        reinterpret_cast<float *>(tensor->ptr), tensor->strides[0], indices->ptr, indices->dtype, indices->strides[0],
        reinterpret_cast<const float *>(value->ptr), value->strides[0], indices->shape[0], rows_t, cols4, col_groups,
        chunks_eff, rpc_r, index_error_flag());
    B200_LAUNCH_CHECK();
    return B200_OK;
  }
  int chunks;
  int64_t rpc;
  pick_chunks(lanes, tensor->shape[dim], chunks, rpc);
  select_add_kernel<<<grid_for(lanes * chunks, 128), 128, 0, resolve_stream(s)>>>(
      to_idx(*tensor), to_idx(*indices), to_idx(*value), dim, lanes, chunks, rpc, index_error_flag());
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int32_t b200_launch_arange(const b200_tensor *out, int64_t start, int64_t step, b200_stream s) {
  B200_REQUIRE(out && out->ptr, B200_ERR_INVALID, "null argument");
  B200_REQUIRE(is_contiguous(*out), B200_ERR_UNSUPPORTED, "arange output must be contiguous");
  B200_REQUIRE(out->dtype == B200_I32 || out->dtype == B200_I64 || out->dtype == B200_F32, B200_ERR_INVALID,
               "arange dtype %d unsupported", out->dtype);
  const int64_t n = numel_of(out->shape, out->rank);
  if (n == 0) return B200_OK;
  arange_kernel<<<grid_for(n, 256), 256, 0, resolve_stream(s)>>>(out->ptr, out->dtype, n, start, step);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int32_t b200_launch_random(const b200_tensor *out, int32_t kind, double lo, double hi, uint64_t seed,
                                      uint64_t offset, b200_stream s) {
  B200_REQUIRE(out && out->ptr, B200_ERR_INVALID, "null argument");
  B200_REQUIRE(is_contiguous(*out), B200_ERR_UNSUPPORTED, "random output must be contiguous");
  B200_REQUIRE(kind >= 0 && kind <= 2, B200_ERR_INVALID, "random kind %d", kind);
  const int64_t n = numel_of(out->shape, out->rank);
  if (n == 0) return B200_OK;
  random_kernel<<<grid_for((n + 3) / 4, 256), 256, 0, resolve_stream(s)>>>(out->ptr, out->dtype, n, kind, (float)lo,
                                                                          (float)hi, seed, offset);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

// ---- data movement composed from strided copies (no kernel of their own but flip):
// float_slice_assign / float_cat / float_repeat_dim / float_flip
// (crates/burn-cubecl/src/kernel/index/{slice_assign,flip,repeat_dim}.rs, crates/burn-backend/src/backend/ops/tensor.rs:161,410,592,1460)
static b200_tensor slice_view(const b200_tensor &t, const int64_t *starts, const int64_t *ends) {
  b200_tensor v = t;
  int64_t off = 0;
  for (int d = 0; d < t.rank; ++d) {
    off += starts[d] * t.strides[d];
    v.shape[d] = ends[d] - starts[d];
  }
  v.ptr = reinterpret_cast<char *>(t.ptr) + off * dtype_size(t.dtype);
  return v;
}

extern "C" int32_t b200_launch_slice_assign(const b200_tensor *tensor, const int64_t *starts, const int64_t *ends,
                                            const b200_tensor *value, b200_stream s) {
  B200_REQUIRE(tensor && starts && ends && value, B200_ERR_INVALID, "null argument");
  B200_REQUIRE(tensor->rank == value->rank, B200_ERR_SHAPE, "slice_assign rank mismatch %d vs %d", tensor->rank, value->rank);
  for (int d = 0; d < tensor->rank; ++d) {
    B200_REQUIRE(starts[d] >= 0 && starts[d] <= ends[d] && ends[d] <= tensor->shape[d], B200_ERR_SHAPE,
                 "slice_assign: range [%lld, %lld) out of bounds for dim %d of size %lld", (long long)starts[d], (long long)ends[d], d,
                 (long long)tensor->shape[d]);
    B200_REQUIRE(value->shape[d] == ends[d] - starts[d], B200_ERR_SHAPE, "slice_assign: value dim %d is %lld, the range holds %lld", d,
                 (long long)value->shape[d], (long long)(ends[d] - starts[d]));
  }
  const b200_tensor view = slice_view(*tensor, starts, ends);
  if (numel_of(view.shape, view.rank) == 0) return B200_OK;
  return b200_launch_copy(value, &view, s);
}

extern "C" int32_t b200_launch_cat(const b200_tensor *inputs, int32_t n, int32_t dim, const b200_tensor *out, b200_stream s) {
  B200_REQUIRE(inputs && out && n >= 1, B200_ERR_INVALID, "null argument / no inputs");
  B200_REQUIRE(dim >= 0 && dim < out->rank, B200_ERR_SHAPE, "cat dim %d out of range", dim);
  int64_t starts[B200_MAX_RANK] = {0}, ends[B200_MAX_RANK];
  int64_t total = 0;
  for (int i = 0; i < n; ++i) {
    B200_REQUIRE(inputs[i].rank == out->rank, B200_ERR_SHAPE, "cat: input %d has rank %d, expected %d", i, inputs[i].rank, out->rank);
    for (int d = 0; d < out->rank; ++d)
      if (d != dim)
        B200_REQUIRE(inputs[i].shape[d] == out->shape[d], B200_ERR_SHAPE, "cat: input %d dim %d is %lld, expected %lld", i, d,
                     (long long)inputs[i].shape[d], (long long)out->shape[d]);
    total += inputs[i].shape[dim];
  }
  B200_REQUIRE(total == out->shape[dim], B200_ERR_SHAPE, "cat: inputs hold %lld along dim %d, the output %lld", (long long)total, dim,
               (long long)out->shape[dim]);
  for (int d = 0; d < out->rank; ++d) ends[d] = out->shape[d];
  int64_t at = 0;
  for (int i = 0; i < n; ++i) {
    if (inputs[i].shape[dim] == 0) continue;         // empty tensors are skipped (cat.rs:92-150)
    starts[dim] = at;
    ends[dim] = at + inputs[i].shape[dim];
    const b200_tensor view = slice_view(*out, starts, ends);
    if (numel_of(view.shape, view.rank) > 0) {
      int32_t st = b200_launch_copy(&inputs[i], &view, s);
      if (st != B200_OK) return st;
    }
    at = ends[dim];
  }
  return B200_OK;
}

extern "C" int32_t b200_launch_repeat_dim(const b200_tensor *input, int32_t dim, int64_t times, const b200_tensor *out, b200_stream s) {
  B200_REQUIRE(input && out, B200_ERR_INVALID, "null argument");
  B200_REQUIRE(input->rank == out->rank && dim >= 0 && dim < input->rank && times >= 0, B200_ERR_SHAPE, "repeat_dim: bad rank / dim / times");
  for (int d = 0; d < input->rank; ++d)
    B200_REQUIRE(out->shape[d] == (d == dim ? input->shape[d] * times : input->shape[d]), B200_ERR_SHAPE, "repeat_dim: bad output dim %d", d);
  int64_t starts[B200_MAX_RANK] = {0}, ends[B200_MAX_RANK];
  for (int d = 0; d < out->rank; ++d) ends[d] = out->shape[d];
  if (numel_of(out->shape, out->rank) == 0) return B200_OK;
  for (int64_t t = 0; t < times; ++t) {
    starts[dim] = t * input->shape[dim];
    ends[dim] = starts[dim] + input->shape[dim];
    const b200_tensor view = slice_view(*out, starts, ends);
    int32_t st = b200_launch_copy(input, &view, s);
    if (st != B200_OK) return st;
  }
  return B200_OK;
}

extern "C" int32_t b200_launch_flip(const b200_tensor *input, const int32_t *axes, int32_t n_axes, const b200_tensor *out, b200_stream s) {
  B200_REQUIRE(input && out && (axes || n_axes == 0), B200_ERR_INVALID, "null argument");
  B200_REQUIRE(input->rank == out->rank && input->dtype == out->dtype, B200_ERR_SHAPE, "flip: rank / dtype mismatch");
  uint32_t mask = 0;
  for (int i = 0; i < n_axes; ++i) {
    B200_REQUIRE(axes[i] >= 0 && axes[i] < input->rank, B200_ERR_SHAPE, "flip axis %d out of range", axes[i]);
    mask |= 1u << axes[i];
  }
  for (int d = 0; d < input->rank; ++d) B200_REQUIRE(input->shape[d] == out->shape[d], B200_ERR_SHAPE, "flip: shape mismatch at dim %d", d);
  const int64_t n = numel_of(out->shape, out->rank);
  if (n == 0) return B200_OK;
  flip_kernel<<<grid_for(n, 256), 256, 0, resolve_stream(s)>>>(to_idx(*input), to_idx(*out), mask, dtype_size(input->dtype), n);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
