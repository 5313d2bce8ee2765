// Row-resident softmax / log_softmax / layer_norm along the last axis — the single-pass
// counterpart of the reference's ReduceBroadcasted optimization
// (crates/burn-cubecl-fusion/src/optim/reduce_broadcasted/unit.rs:65), which fuses the chains
//   softmax     = max_dim, sub, exp, sum_dim, div            (crates/burn-backend/src/backend/ops/activation.rs:250-256)
//   log_softmax = max_dim, sub, exp, sum_dim, log, sub       (activation.rs:271-276)
//   layer_norm  = mean_dim, sub, mul, mean_dim, add_scalar, sqrt, div, mul, add
//                                                            (crates/burn-backend/src/backend/ops/modules/base.rs:846-877)
// into one kernel.  Each element goes through exactly the f32 operations of the op-by-op chain
// (IEEE sub/div/sqrt, expf/logf); only the order of the row sums differs from the oracle
// (<= 1e-5 relative, like every reduction here).
//
// One row stays on chip between the passes: a warp keeps a row of up to 2048 elements in
// registers (128-bit loads, shuffle trees); longer rows (up to ~57 K f32, e.g. a 50 257-entry
// vocabulary) are staged once in shared memory by a whole CTA.  HBM traffic = one read + one
// write of the tensor.  Roofline: HBM; algorithmic bytes = 8 B / element.
#include <cooperative_groups.h>
#include <cstdlib>

#include "common.cuh"

namespace b200 {
namespace rn {

constexpr int kBlock = 256;
constexpr int kWarps = kBlock / 32;

enum Mode : int { kSoftmax = 0, kLogSoftmax = 1, kLayerNorm = 2 };

struct Params {
  const float *x;
  float *y;
  const float *gamma;  // layer_norm scale (may be null)
  const float *beta;   // layer_norm shift (may be null)
  int64_t x_row_stride, y_row_stride;
  uint32_t rows, R;
  float eps;
  int32_t vec;   // CTA kernel: rows are 16-byte aligned and R % 4 == 0
};

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, m));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, m));
  return v;
}
// NaN-propagating max like the oracle's max_dim (first NaN wins → NaN)
__device__ __forceinline__ float nan_max(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fmaxf(a, b); }

// ---- a warp owns a row; V float4 per lane live in registers
template <int MODE, int V>
__global__ void __launch_bounds__(kBlock) row_warp_kernel(const Params P) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t r4 = P.R >> 2;
  for (uint32_t row = blockIdx.x * kWarps + warp; row < P.rows; row += gridDim.x * kWarps) {
    const float4 *xr = reinterpret_cast<const float4 *>(P.x + (int64_t)row * P.x_row_stride);
    float4 *yr = reinterpret_cast<float4 *>(P.y + (int64_t)row * P.y_row_stride);
    float4 v[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const uint32_t i = lane + k * 32;
      v[k] = i < r4 ? __ldcs(xr + i) : make_float4(0, 0, 0, 0);
    }
    if constexpr (MODE == kLayerNorm) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < V; ++k)
        if (lane + k * 32 < r4) s = __fadd_rn(s, __fadd_rn(__fadd_rn(v[k].x, v[k].y), __fadd_rn(v[k].z, v[k].w)));
      const float mean = __fdiv_rn(warp_sum(s), (float)P.R);
      float q = 0.f;
#pragma unroll
      for (int k = 0; k < V; ++k) {
        v[k].x = __fsub_rn(v[k].x, mean); v[k].y = __fsub_rn(v[k].y, mean);
        v[k].z = __fsub_rn(v[k].z, mean); v[k].w = __fsub_rn(v[k].w, mean);
        if (lane + k * 32 < r4)
          q = __fadd_rn(q, __fadd_rn(__fadd_rn(__fmul_rn(v[k].x, v[k].x), __fmul_rn(v[k].y, v[k].y)),
                                     __fadd_rn(__fmul_rn(v[k].z, v[k].z), __fmul_rn(v[k].w, v[k].w))));
      }
      const float var = __fdiv_rn(warp_sum(q), (float)P.R);
      const float denom = __fsqrt_rn(__fadd_rn(var, P.eps));
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const uint32_t i = lane + k * 32;
        if (i >= r4) continue;
        float4 o;
        o.x = __fdiv_rn(v[k].x, denom); o.y = __fdiv_rn(v[k].y, denom);
        o.z = __fdiv_rn(v[k].z, denom); o.w = __fdiv_rn(v[k].w, denom);
        if (P.gamma) {
          const float4 g = __ldg(reinterpret_cast<const float4 *>(P.gamma) + i);
          o.x = __fmul_rn(o.x, g.x); o.y = __fmul_rn(o.y, g.y); o.z = __fmul_rn(o.z, g.z); o.w = __fmul_rn(o.w, g.w);
        }
        if (P.beta) {
          const float4 b = __ldg(reinterpret_cast<const float4 *>(P.beta) + i);
          o.x = __fadd_rn(o.x, b.x); o.y = __fadd_rn(o.y, b.y); o.z = __fadd_rn(o.z, b.z); o.w = __fadd_rn(o.w, b.w);
        }
        __stcs(yr + i, o);
      }
    } else {
      float m = -INFINITY;
#pragma unroll
      for (int k = 0; k < V; ++k)
        if (lane + k * 32 < r4) m = nan_max(m, nan_max(nan_max(v[k].x, v[k].y), nan_max(v[k].z, v[k].w)));
      {  // NaN-aware warp max
#pragma unroll
        for (int s = 16; s >= 1; s >>= 1) m = nan_max(m, __shfl_xor_sync(0xffffffffu, m, s));
      }
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < V; ++k) {
        v[k].x = __fsub_rn(v[k].x, m); v[k].y = __fsub_rn(v[k].y, m);
        v[k].z = __fsub_rn(v[k].z, m); v[k].w = __fsub_rn(v[k].w, m);
        float4 e;
        e.x = expf(v[k].x); e.y = expf(v[k].y); e.z = expf(v[k].z); e.w = expf(v[k].w);
        if (lane + k * 32 < r4) s = __fadd_rn(s, __fadd_rn(__fadd_rn(e.x, e.y), __fadd_rn(e.z, e.w)));
        if constexpr (MODE == kSoftmax) v[k] = e;
      }
      s = warp_sum(s);
      const float ls = logf(s);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        const uint32_t i = lane + k * 32;
        if (i >= r4) continue;
        float4 o;
        if constexpr (MODE == kSoftmax) {
          o.x = __fdiv_rn(v[k].x, s); o.y = __fdiv_rn(v[k].y, s); o.z = __fdiv_rn(v[k].z, s); o.w = __fdiv_rn(v[k].w, s);
        } else {
          o.x = __fsub_rn(v[k].x, ls); o.y = __fsub_rn(v[k].y, ls); o.z = __fsub_rn(v[k].z, ls); o.w = __fsub_rn(v[k].w, ls);
        }
        __stcs(yr + i, o);
      }
    }
  }
}

// ---- a CTA owns a row staged in shared memory (any R that fits, scalar accesses allowed)
__device__ __forceinline__ float block_reduce_sum(float v, float *scratch) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = (threadIdx.x & 31) < kWarps ? scratch[threadIdx.x & 31] : 0.f;
  return warp_sum(r);
}
__device__ __forceinline__ float block_reduce_max(float v, float *scratch) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) v = nan_max(v, __shfl_xor_sync(0xffffffffu, v, s));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = (threadIdx.x & 31) < kWarps ? scratch[threadIdx.x & 31] : -INFINITY;
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) r = nan_max(r, __shfl_xor_sync(0xffffffffu, r, s));
  return r;
}

template <int MODE>
__global__ void __launch_bounds__(kBlock) row_cta_kernel(const Params P) {
  extern __shared__ __align__(16) float rowbuf[];
  __shared__ float scratch[kWarps];
  for (uint32_t row = blockIdx.x; row < P.rows; row += gridDim.x) {
    const float *xr = P.x + (int64_t)row * P.x_row_stride;
    float *yr = P.y + (int64_t)row * P.y_row_stride;
    __syncthreads();
    if (P.vec) {   // 128-bit loads when the row is 16-byte aligned and R % 4 == 0
      const float4 *x4 = reinterpret_cast<const float4 *>(xr);
      float4 *b4 = reinterpret_cast<float4 *>(rowbuf);
      for (uint32_t i = threadIdx.x; i < (P.R >> 2); i += kBlock) b4[i] = __ldcs(x4 + i);
    } else {
      for (uint32_t i = threadIdx.x; i < P.R; i += kBlock) rowbuf[i] = __ldcs(xr + i);
    }
    __syncthreads();
    if constexpr (MODE == kLayerNorm) {
      float s = 0.f;
      for (uint32_t i = threadIdx.x; i < P.R; i += kBlock) s = __fadd_rn(s, rowbuf[i]);
      const float mean = __fdiv_rn(block_reduce_sum(s, scratch), (float)P.R);
      float q = 0.f;
      for (uint32_t i = threadIdx.x; i < P.R; i += kBlock) {
        const float c = __fsub_rn(rowbuf[i], mean);
        rowbuf[i] = c;
        q = __fadd_rn(q, __fmul_rn(c, c));
      }
      const float var = __fdiv_rn(block_reduce_sum(q, scratch), (float)P.R);
      const float denom = __fsqrt_rn(__fadd_rn(var, P.eps));
      for (uint32_t i = threadIdx.x; i < P.R; i += kBlock) {
        float o = __fdiv_rn(rowbuf[i], denom);
        if (P.gamma) o = __fmul_rn(o, __ldg(P.gamma + i));
        if (P.beta) o = __fadd_rn(o, __ldg(P.beta + i));
        __stcs(yr + i, o);
      }
    } else {
      float m = -INFINITY;
      for (uint32_t i = threadIdx.x; i < P.R; i += kBlock) m = nan_max(m, rowbuf[i]);
      m = block_reduce_max(m, scratch);
      float s = 0.f;
      for (uint32_t i = threadIdx.x; i < P.R; i += kBlock) {
        const float sh = __fsub_rn(rowbuf[i], m);
        const float e = expf(sh);
        s = __fadd_rn(s, e);
        rowbuf[i] = MODE == kSoftmax ? e : sh;
      }
      s = block_reduce_sum(s, scratch);
      const float ls = logf(s);
      if (P.vec) {
        const float4 *b4 = reinterpret_cast<const float4 *>(rowbuf);
        float4 *y4 = reinterpret_cast<float4 *>(yr);
        for (uint32_t i = threadIdx.x; i < (P.R >> 2); i += kBlock) {
          float4 o = b4[i];
          if (MODE == kSoftmax) { o.x = __fdiv_rn(o.x, s); o.y = __fdiv_rn(o.y, s); o.z = __fdiv_rn(o.z, s); o.w = __fdiv_rn(o.w, s); }
          else { o.x = __fsub_rn(o.x, ls); o.y = __fsub_rn(o.y, ls); o.z = __fsub_rn(o.z, ls); o.w = __fsub_rn(o.w, ls); }
          __stcs(y4 + i, o);
        }
      } else {
        for (uint32_t i = threadIdx.x; i < P.R; i += kBlock)
          __stcs(yr + i, MODE == kSoftmax ? __fdiv_rn(rowbuf[i], s) : __fsub_rn(rowbuf[i], ls));
      }
    }
  }
}

// Collapses [..., R] into (rows, row stride); requires a unit inner stride and leading dims that
// are jointly strided.
static int32_t rows_view(const b200_tensor &t, uint32_t &rows, uint32_t &R, int64_t &row_stride, const char *what) {
  B200_REQUIRE(t.rank >= 1 && t.rank <= B200_MAX_RANK, B200_ERR_INVALID, "%s: bad rank %d", what, t.rank);
  B200_REQUIRE(t.dtype == B200_F32, B200_ERR_UNSUPPORTED, "%s: only f32 is implemented", what);
  const int last = t.rank - 1;
  B200_REQUIRE(t.shape[last] == 1 || t.strides[last] == 1, B200_ERR_UNSUPPORTED, "%s: the normalised axis must be contiguous", what);
  int64_t n_rows = 1, stride = t.shape[last];
  bool first = true;
  for (int d = last - 1; d >= 0; --d) {
    if (t.shape[d] == 1) continue;
    if (first) { stride = t.strides[d]; first = false; }
    B200_REQUIRE(t.strides[d] == stride * n_rows, B200_ERR_UNSUPPORTED, "%s: leading dims must be jointly strided", what);
    n_rows *= t.shape[d];
  }
  B200_REQUIRE(n_rows < (1ll << 31) && t.shape[last] < (1ll << 31), B200_ERR_UNSUPPORTED, "%s: too large", what);
  rows = (uint32_t)n_rows;
  R = (uint32_t)t.shape[last];
  row_stride = stride;
  return B200_OK;
}

template <int MODE>
static int32_t launch_rows(const Params &P, cudaStream_t stream) {
  if (P.rows == 0 || P.R == 0) return B200_OK;
  const bool vec_ok = P.R % 4 == 0 && ((uintptr_t)P.x) % 16 == 0 && ((uintptr_t)P.y) % 16 == 0 &&
                      P.x_row_stride % 4 == 0 && P.y_row_stride % 4 == 0 &&
                      (!P.gamma || ((uintptr_t)P.gamma) % 16 == 0) && (!P.beta || ((uintptr_t)P.beta) % 16 == 0);
  const int sms = sm_count();
  if (vec_ok && P.R <= 2048) {
    const uint32_t r4 = P.R / 4;
    const unsigned grid = (unsigned)std::min<uint32_t>((P.rows + kWarps - 1) / kWarps, (uint32_t)sms * 8u);
    if (r4 <= 32) row_warp_kernel<MODE, 1><<<grid, kBlock, 0, stream>>>(P);
    else if (r4 <= 64) row_warp_kernel<MODE, 2><<<grid, kBlock, 0, stream>>>(P);
    else if (r4 <= 128) row_warp_kernel<MODE, 4><<<grid, kBlock, 0, stream>>>(P);
    else if (r4 <= 256) row_warp_kernel<MODE, 8><<<grid, kBlock, 0, stream>>>(P);
    else row_warp_kernel<MODE, 16><<<grid, kBlock, 0, stream>>>(P);
    B200_LAUNCH_CHECK();
    return B200_OK;
  }
  const size_t smem = (size_t)P.R * 4;
  B200_REQUIRE((int)smem + 1024 <= max_smem_optin(), B200_ERR_UNSUPPORTED,
               "row of %u f32 does not fit in shared memory; use the op chain", P.R);
  Params Q = P;
  Q.vec = vec_ok ? 1 : 0;
  auto kern = row_cta_kernel<MODE>;
  if (smem > 48 * 1024) B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kBlock, smem));
  const unsigned grid = (unsigned)std::min<uint32_t>(P.rows, (uint32_t)(sms * std::max(per_sm, 1)));
  kern<<<grid, kBlock, smem, stream>>>(Q);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

// ---- softmax cross-entropy, forward value and logits gradient in one pass over a row staged in smem:
//   picked[r] = log_softmax(x[r])[t[r]]          (what log_softmax(..).gather(1, targets) yields,
//                                                  crates/burn-nn/src/loss/cross_entropy.rs:171-181)
//   dx[r, i]  = (softmax(x[r])[i] - [i == t[r]]) * grad_scale      (its gradient under mean().neg())
// dx may alias x.  HBM traffic: read the logits once, write the gradient once (8 B/elem) instead of
// log_softmax (8) + gather + the exp/one-hot tape (8+).
struct XentParams {
  const float *x;
  float *dx;
  const void *targets;
  float *picked;
  int64_t x_stride, dx_stride;
  uint32_t rows, R;
  int32_t target_i64, vec;
  float grad_scale;
  int32_t *err;     // host-mapped sticky flag: out-of-range target
};

// 1024 threads: a 200 KB vocabulary row leaves room for one CTA per SM, so the CTA itself has to keep
// enough 128-bit loads in flight
constexpr int kXBlock = 1024;
__device__ __forceinline__ float xblock_sum(float v, float *scratch) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  return warp_sum(scratch[threadIdx.x & 31]);          // kXBlock / 32 == 32 partials
}
__device__ __forceinline__ float xblock_max(float v, float *scratch) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) v = nan_max(v, __shfl_xor_sync(0xffffffffu, v, s));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = scratch[threadIdx.x & 31];
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) r = nan_max(r, __shfl_xor_sync(0xffffffffu, r, s));
  return r;
}
__device__ __forceinline__ float xent_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void xent_cp16(void *smem_dst, const void *g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(g) : "memory");
}
// Pipelining: the row lives in smem (one CTA per SM for a 200 KB vocabulary row), so a CTA that loads, reduces
// and writes strictly in turn leaves HBM idle two thirds of the time (0.47 of peak).  Here the gradient pass,
// which frees smem slot i the moment it has read it, refills that slot with element i of the CTA's NEXT row
// through cp.async: the next row's read streams in while the current row's gradient streams out, and the
// max / sum passes between them touch shared memory only.
__global__ void __launch_bounds__(kXBlock) softmax_xent_kernel(const XentParams P) {
  extern __shared__ __align__(16) float rowbuf[];
  __shared__ float scratch[kXBlock / 32];
  const uint32_t R4 = P.R >> 2;
  if (P.vec && blockIdx.x < P.rows) {   // prologue: this CTA's first row
    const float4 *x4 = reinterpret_cast<const float4 *>(P.x + (int64_t)blockIdx.x * P.x_stride);
    float4 *b4 = reinterpret_cast<float4 *>(rowbuf);
    for (uint32_t i = threadIdx.x; i < R4; i += kXBlock) xent_cp16(b4 + i, x4 + i);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  }
  for (uint32_t row = blockIdx.x; row < P.rows; row += gridDim.x) {
    const float *xr = P.x + (int64_t)row * P.x_stride;
    float *dr = P.dx + (int64_t)row * P.dx_stride;
    int64_t t = P.target_i64 ? reinterpret_cast<const long long *>(P.targets)[row]
                             : (int64_t) reinterpret_cast<const int32_t *>(P.targets)[row];
    if (t < 0 || t >= (int64_t)P.R) {     // the reference's gather would panic: flag it, contribute 0 to the loss
      if (threadIdx.x == 0) *P.err = kIdxErrTarget;
      t = -1;
    }
    if (P.vec) {
      asm volatile("cp.async.wait_all;\n" ::: "memory");
    } else {
      __syncthreads();
      for (uint32_t i = threadIdx.x; i < P.R; i += kXBlock) rowbuf[i] = __ldcs(xr + i);
    }
    __syncthreads();
    // The kernel was as much issue-bound as memory-bound: two libm expf per element (~20 instructions each) on a
    // 50k-element row is ~8 us of issue time per row per SM against 9 us of HBM time.  exp runs in the base-2
    // domain instead (one FFMA + one MUFU.EX2 per element, 2 ulp + |x|*2^-24 relative: <= 2e-6 on a softmax
    // value, inside the 1e-5 reduction tolerance); the row log-sum-exp itself keeps libm's logf.
    const float kL2e = 1.4426950408889634f;
    float m = -INFINITY;
    for (uint32_t i = threadIdx.x; i < P.R; i += kXBlock) m = nan_max(m, rowbuf[i]);
    m = xblock_max(m, scratch);
    const float mb = __fmul_rn(m, kL2e);
    float s = 0.f;
    for (uint32_t i = threadIdx.x; i < P.R; i += kXBlock) s = __fadd_rn(s, xent_ex2(fmaf(rowbuf[i], kL2e, -mb)));
    s = xblock_sum(s, scratch);
    const float ls = logf(s);
    if (threadIdx.x == 0) P.picked[row] = t >= 0 ? __fsub_rn(__fsub_rn(rowbuf[t], m), ls) : 0.0f;
    // gradient: softmax - onehot, scaled: softmax = 2^(x*log2e - (m*log2e + log2 s))
    const float shift = __fadd_rn(mb, __log2f(s));
    if (P.vec) {
      __syncthreads();                     // rowbuf[t] has been read before any slot is refilled
      float4 *b4 = reinterpret_cast<float4 *>(rowbuf);
      float4 *d4 = reinterpret_cast<float4 *>(dr);
      const uint32_t next = row + gridDim.x;
      const float4 *n4 = reinterpret_cast<const float4 *>(P.x + (int64_t)next * P.x_stride);
      const bool more = next < P.rows;
      for (uint32_t i = threadIdx.x; i < R4; i += kXBlock) {
        const float4 xv = b4[i];
        if (more) xent_cp16(b4 + i, n4 + i);   // this slot is free: prefetch the next row's element
        float4 o;
        o.x = xent_ex2(fmaf(xv.x, kL2e, -shift)); o.y = xent_ex2(fmaf(xv.y, kL2e, -shift));
        o.z = xent_ex2(fmaf(xv.z, kL2e, -shift)); o.w = xent_ex2(fmaf(xv.w, kL2e, -shift));
        const int64_t c = (int64_t)i * 4;
        if (t >= c && t < c + 4) {
          if (t == c) o.x = __fsub_rn(o.x, 1.0f);
          else if (t == c + 1) o.y = __fsub_rn(o.y, 1.0f);
          else if (t == c + 2) o.z = __fsub_rn(o.z, 1.0f);
          else o.w = __fsub_rn(o.w, 1.0f);
        }
        o.x = __fmul_rn(o.x, P.grad_scale); o.y = __fmul_rn(o.y, P.grad_scale);
        o.z = __fmul_rn(o.z, P.grad_scale); o.w = __fmul_rn(o.w, P.grad_scale);
        __stcs(d4 + i, o);
      }
      asm volatile("cp.async.commit_group;\n" ::: "memory");
    } else {
      for (uint32_t i = threadIdx.x; i < P.R; i += kXBlock) {
        float o = xent_ex2(fmaf(rowbuf[i], kL2e, -shift));
        if ((int64_t)i == t) o = __fsub_rn(o, 1.0f);
        __stcs(dr + i, __fmul_rn(o, P.grad_scale));
      }
    }
  }
}

// The same computation for rows too long for two CTAs per SM (a 50 260-class row is 196 KB): a CLUSTER of two CTAs
// owns a row, each staging one half (98 KB, 512 threads), so every SM hosts halves of two different rows whose phases
// drift apart — one streams its gradient out / its next row in while the other runs the shared-memory-only max and
// sum passes.  The halves meet twice per row through distributed shared memory (max, then sum-of-exponentials, added
// rank 0 first so the result does not depend on which CTA asks).  Exchange slots are double-buffered by row parity:
// a slot is rewritten two rows later, after a cluster barrier the peer can only have passed once it had read it.
constexpr int kXPair = 512;
__device__ __forceinline__ float pblock_sum(float v, float *scratch) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  return warp_sum((threadIdx.x & 31) < kXPair / 32 ? scratch[threadIdx.x & 31] : 0.0f);
}
__device__ __forceinline__ float pblock_max(float v, float *scratch) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) v = nan_max(v, __shfl_xor_sync(0xffffffffu, v, s));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = (threadIdx.x & 31) < kXPair / 32 ? scratch[threadIdx.x & 31] : -INFINITY;
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) r = nan_max(r, __shfl_xor_sync(0xffffffffu, r, s));
  return r;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kXPair, 2) softmax_xent_pair_kernel(const XentParams P) {
  namespace cg = cooperative_groups;
  extern __shared__ __align__(16) float rowbuf[];   // this CTA's half of the row
  __shared__ float scratch[kXPair / 32];
  __shared__ float xch[2][2];                        // [row parity][0 = max, 1 = sum]
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t crank = cluster.block_rank();
  const float *peer = cluster.map_shared_rank(&xch[0][0], crank ^ 1u);
  const uint32_t pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const uint32_t R4 = P.R >> 2, h4 = (R4 + 1) >> 1;
  const uint32_t lo4 = crank * h4, hi4 = min(R4, lo4 + h4), n4 = hi4 > lo4 ? hi4 - lo4 : 0u;
  float4 *b4 = reinterpret_cast<float4 *>(rowbuf);
  if (pair < P.rows) {
    const float4 *x4 = reinterpret_cast<const float4 *>(P.x + (int64_t)pair * P.x_stride) + lo4;
    for (uint32_t i = threadIdx.x; i < n4; i += kXPair) xent_cp16(b4 + i, x4 + i);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  }
  const float kL2e = 1.4426950408889634f;
  uint32_t parity = 0;
  for (uint32_t row = pair; row < P.rows; row += n_pairs, parity ^= 1u) {
    float *dr = P.dx + (int64_t)row * P.dx_stride;
    int64_t t = P.target_i64 ? reinterpret_cast<const long long *>(P.targets)[row]
                             : (int64_t) reinterpret_cast<const int32_t *>(P.targets)[row];
    if (t < 0 || t >= (int64_t)P.R) {
      if (threadIdx.x == 0 && crank == 0) *P.err = kIdxErrTarget;
      t = -1;
    }
    asm volatile("cp.async.wait_all;\n" ::: "memory");
    __syncthreads();
    float m = -INFINITY;
    for (uint32_t i = threadIdx.x; i < n4; i += kXPair) {
      const float4 v = b4[i];
      m = nan_max(m, nan_max(nan_max(v.x, v.y), nan_max(v.z, v.w)));
    }
    m = pblock_max(m, scratch);
    if (threadIdx.x == 0) xch[parity][0] = m;
    cluster.sync();
    m = nan_max(m, peer[parity * 2 + 0]);
    const float mb = __fmul_rn(m, kL2e);
    float s = 0.f;
    for (uint32_t i = threadIdx.x; i < n4; i += kXPair) {
      const float4 v = b4[i];
      s = __fadd_rn(s, __fadd_rn(__fadd_rn(xent_ex2(fmaf(v.x, kL2e, -mb)), xent_ex2(fmaf(v.y, kL2e, -mb))),
                                 __fadd_rn(xent_ex2(fmaf(v.z, kL2e, -mb)), xent_ex2(fmaf(v.w, kL2e, -mb)))));
    }
    s = pblock_sum(s, scratch);
    if (threadIdx.x == 0) xch[parity][1] = s;
    cluster.sync();
    const float other = peer[parity * 2 + 1];
    s = crank == 0 ? __fadd_rn(s, other) : __fadd_rn(other, s);      // rank 0's half first, on both CTAs
    const float ls = logf(s);
    const int64_t tl = t - (int64_t)lo4 * 4;                          // the target's position inside this half
    const bool mine = t >= 0 && tl >= 0 && tl < (int64_t)n4 * 4;
    if (threadIdx.x == 0) {
      if (mine) P.picked[row] = __fsub_rn(__fsub_rn(rowbuf[tl], m), ls);
      else if (t < 0 && crank == 0) P.picked[row] = 0.0f;
    }
    const float shift = __fadd_rn(mb, __log2f(s));
    __syncthreads();                       // rowbuf[tl] has been read before any slot is refilled
    float4 *d4 = reinterpret_cast<float4 *>(dr) + lo4;
    const uint32_t next = row + n_pairs;
    const float4 *n4p = reinterpret_cast<const float4 *>(P.x + (int64_t)next * P.x_stride) + lo4;
    const bool more = next < P.rows;
    for (uint32_t i = threadIdx.x; i < n4; i += kXPair) {
      const float4 xv = b4[i];
      if (more) xent_cp16(b4 + i, n4p + i);
      float4 o;
      o.x = xent_ex2(fmaf(xv.x, kL2e, -shift)); o.y = xent_ex2(fmaf(xv.y, kL2e, -shift));
      o.z = xent_ex2(fmaf(xv.z, kL2e, -shift)); o.w = xent_ex2(fmaf(xv.w, kL2e, -shift));
      const int64_t c = (int64_t)i * 4;
      if (mine && tl >= c && tl < c + 4) {
        if (tl == c) o.x = __fsub_rn(o.x, 1.0f);
        else if (tl == c + 1) o.y = __fsub_rn(o.y, 1.0f);
        else if (tl == c + 2) o.z = __fsub_rn(o.z, 1.0f);
        else o.w = __fsub_rn(o.w, 1.0f);
      }
      o.x = __fmul_rn(o.x, P.grad_scale); o.y = __fmul_rn(o.y, P.grad_scale);
      o.z = __fmul_rn(o.z, P.grad_scale); o.w = __fmul_rn(o.w, P.grad_scale);
      __stcs(d4 + i, o);
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  }
  cluster.sync();                          // no CTA may exit while its peer can still read its exchange slots
}

}  // namespace rn
}  // namespace b200

using namespace b200;

extern "C" int32_t b200_launch_softmax(const b200_tensor *input, const b200_tensor *out, int32_t log_softmax,
                                       b200_stream s) {
  B200_REQUIRE(input && out && input->ptr && out->ptr, B200_ERR_INVALID, "null argument");
  B200_REQUIRE(input->rank == out->rank, B200_ERR_SHAPE, "softmax rank mismatch");
  for (int d = 0; d < input->rank; ++d)
    B200_REQUIRE(input->shape[d] == out->shape[d], B200_ERR_SHAPE, "softmax shape mismatch at dim %d", d);
  rn::Params P;
  memset(&P, 0, sizeof(P));
  uint32_t rows2, R2;
  int32_t st = rn::rows_view(*input, P.rows, P.R, P.x_row_stride, "softmax input");
  if (st != B200_OK) return st;
  st = rn::rows_view(*out, rows2, R2, P.y_row_stride, "softmax output");
  if (st != B200_OK) return st;
  P.x = reinterpret_cast<const float *>(input->ptr);
  P.y = reinterpret_cast<float *>(out->ptr);
  return log_softmax ? rn::launch_rows<rn::kLogSoftmax>(P, resolve_stream(s))
                     : rn::launch_rows<rn::kSoftmax>(P, resolve_stream(s));
}

extern "C" int32_t b200_launch_layer_norm(const b200_tensor *input, const b200_tensor *gamma, const b200_tensor *beta,
                                          double eps, const b200_tensor *out, b200_stream s) {
  B200_REQUIRE(input && out && input->ptr && out->ptr, B200_ERR_INVALID, "null argument");
  B200_REQUIRE(input->rank == out->rank, B200_ERR_SHAPE, "layer_norm rank mismatch");
  for (int d = 0; d < input->rank; ++d)
    B200_REQUIRE(input->shape[d] == out->shape[d], B200_ERR_SHAPE, "layer_norm shape mismatch at dim %d", d);
  rn::Params P;
  memset(&P, 0, sizeof(P));
  uint32_t rows2, R2;
  int32_t st = rn::rows_view(*input, P.rows, P.R, P.x_row_stride, "layer_norm input");
  if (st != B200_OK) return st;
  st = rn::rows_view(*out, rows2, R2, P.y_row_stride, "layer_norm output");
  if (st != B200_OK) return st;
  auto vec1d = [&](const b200_tensor *t, const float *&dst, const char *what) -> int32_t {
    dst = nullptr;
    if (!t) return B200_OK;
    B200_REQUIRE(t->dtype == B200_F32 && t->ptr, B200_ERR_UNSUPPORTED, "%s must be f32", what);
    int64_t n = 1;
    for (int d = 0; d < t->rank; ++d) n *= t->shape[d];
    B200_REQUIRE(n == (int64_t)P.R && (t->shape[t->rank - 1] == 1 || t->strides[t->rank - 1] == 1), B200_ERR_SHAPE,
                 "%s must hold d_model = %u contiguous values", what, P.R);
    dst = reinterpret_cast<const float *>(t->ptr);
    return B200_OK;
  };
  if ((st = vec1d(gamma, P.gamma, "gamma")) != B200_OK) return st;
  if ((st = vec1d(beta, P.beta, "beta")) != B200_OK) return st;
  P.x = reinterpret_cast<const float *>(input->ptr);
  P.y = reinterpret_cast<float *>(out->ptr);
  P.eps = (float)eps;
  return rn::launch_rows<rn::kLayerNorm>(P, resolve_stream(s));
}

extern "C" int32_t b200_launch_softmax_cross_entropy(const b200_tensor *logits, const b200_tensor *targets, double grad_scale,
                                                     const b200_tensor *picked, const b200_tensor *dlogits, b200_stream s) {
  B200_REQUIRE(logits && targets && picked && dlogits, B200_ERR_INVALID, "null argument");
  B200_REQUIRE(logits->rank == 2 && dlogits->rank == 2 && logits->dtype == B200_F32 && dlogits->dtype == B200_F32 &&
                   logits->ptr && dlogits->ptr,
               B200_ERR_UNSUPPORTED, "softmax_cross_entropy takes f32 [N, V] logits");
  const int64_t N = logits->shape[0], V = logits->shape[1];
  B200_REQUIRE(dlogits->shape[0] == N && dlogits->shape[1] == V, B200_ERR_SHAPE, "gradient shape mismatch");
  B200_REQUIRE((logits->strides[1] == 1 || V == 1) && (dlogits->strides[1] == 1 || V == 1), B200_ERR_UNSUPPORTED,
               "the class axis must be contiguous");
  B200_REQUIRE((targets->dtype == B200_I32 || targets->dtype == B200_I64) && targets->ptr && is_contiguous(*targets),
               B200_ERR_INVALID, "targets must be contiguous i32 / i64");
  int64_t nt = 1, np_ = 1;
  for (int d = 0; d < targets->rank; ++d) nt *= targets->shape[d];
  for (int d = 0; d < picked->rank; ++d) np_ *= picked->shape[d];
  B200_REQUIRE(nt == N && np_ == N && picked->dtype == B200_F32 && picked->ptr && is_contiguous(*picked), B200_ERR_SHAPE,
               "targets / picked must hold N = %lld entries", (long long)N);
  if (N == 0 || V == 0) return B200_OK;
  B200_REQUIRE(N < (1ll << 31) && V < (1ll << 31), B200_ERR_UNSUPPORTED, "softmax_cross_entropy: too large");
  const size_t smem = (size_t)V * 4;
  B200_REQUIRE((int)smem + 1024 <= max_smem_optin(), B200_ERR_UNSUPPORTED,
               "a row of %lld classes does not fit in shared memory; use the op chain", (long long)V);
  rn::XentParams P;
  memset(&P, 0, sizeof(P));
  P.x = reinterpret_cast<const float *>(logits->ptr);
  P.dx = reinterpret_cast<float *>(dlogits->ptr);
  P.targets = targets->ptr;
  P.picked = reinterpret_cast<float *>(picked->ptr);
  P.x_stride = logits->strides[0];
  P.dx_stride = dlogits->strides[0];
  P.rows = (uint32_t)N;
  P.R = (uint32_t)V;
  P.target_i64 = targets->dtype == B200_I64;
  P.vec = V % 4 == 0 && ((uintptr_t)P.x % 16) == 0 && ((uintptr_t)P.dx % 16) == 0 && P.x_stride % 4 == 0 && P.dx_stride % 4 == 0;
  P.grad_scale = (float)grad_scale;
  P.err = index_error_flag();
  auto kern = rn::softmax_xent_kernel;
  if (smem > 48 * 1024) B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, rn::kXBlock, smem));
  static const bool pair_ok = [] { const char *e = std::getenv("B200_XENT_NO_PAIR"); return !(e && e[0] == '1'); }();
  if (per_sm <= 1 && P.vec && V >= 8 && N >= 2 && pair_ok) {
    // one full-row CTA per SM leaves HBM idle during the smem-only passes: a 2-CTA cluster per row, two half rows per SM
    const size_t half = (size_t)(((V / 4) + 1) / 2) * 16;
    auto pk = rn::softmax_xent_pair_kernel;
    int32_t st = ensure_dyn_smem((const void *)pk, half);
    if (st != B200_OK) return st;
    int pair_per_sm = 0;
    B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pair_per_sm, pk, rn::kXPair, half));
    if (pair_per_sm >= 2) {
      const int64_t pairs = std::min<int64_t>(N, (int64_t)sm_count() * pair_per_sm / 2);
      pk<<<(unsigned)(2 * pairs), rn::kXPair, half, resolve_stream(s)>>>(P);
      B200_LAUNCH_CHECK();
      return B200_OK;
    }
  }
  const unsigned grid = (unsigned)std::min<int64_t>(N, (int64_t)sm_count() * std::max(per_sm, 1));
  kern<<<grid, rn::kXBlock, smem, resolve_stream(s)>>>(P);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
