// Row-resident backward kernels for softmax and layer_norm along the last axis — what
// burn-autodiff's reverse pass over the op chains of rownorm.cu collapses to when the row stays
// on chip (the ReduceBroadcasted analogue, backward direction).
//
//   softmax backward   dx = (dy - sum(dy*y)) * y / div          [mask ? 0 : .]
//       the chain rule through max_dim/sub/exp/sum_dim/div (crates/burn-backend/src/backend/ops/activation.rs:250-256)
//       as burn-autodiff accumulates it (crates/burn-autodiff/src/ops/tensor.rs: div/exp/sub/sum_dim backward);
//       `div` and the mask fold in the attention-score scaling and mask_fill backward that follow it
//       in MultiHeadAttention (crates/burn-nn/src/modules/attention/mha.rs:253-311).
//   layer_norm backward  with xh = (x-mean)/sqrt(var+eps), g = dy*gamma:
//       dx = (g - mean(g) - xh*mean(g*xh)) / sqrt(var+eps);  dgamma = sum_rows(dy*xh);  dbeta = sum_rows(dy)
//       (reverse of crates/burn-backend/src/backend/ops/modules/base.rs:846-877).
//       dgamma / dbeta leave the kernel as per-CTA partial rows [G, R]; the caller finishes them with
//       the deterministic column reduce (b200_launch_reduce), so no atomics and no run-to-run noise.
//
// A warp keeps a row of up to 2048 f32 of BOTH operands in registers (128-bit loads, shuffle
// trees); other rows are handled by a CTA that streams the row twice (second pass from L2).
// HBM traffic = read 2 tensors + write 1: 12 B / element.  Roofline: HBM.  Sums are f32 in a
// different order than the op-by-op chain (<= 1e-5 relative, like every reduction here).
#include "common.cuh"

namespace b200 {
namespace rnb {

constexpr int kBlock = 256;
constexpr int kWarps = kBlock / 32;

struct SoftmaxBwd {
  const float *y, *dy;
  float *dx;
  const uint8_t *mask;       // optional: nonzero = masked (dx = 0); row r uses mask row r % mask_rows
  int64_t y_stride, dy_stride, dx_stride, mask_stride;
  uint32_t rows, R, mask_rows;
  float div;
  float mul;                 // 1/div when that is exact (div a power of two): x*mul == x/div bit for bit; else 0
};

struct LayerNormBwd {
  const float *x, *dy, *gamma;
  float *dx, *pgamma, *pbeta;  // partials [gridDim.x, R]
  float *pdx;                  // optional third partial: column sums of dx (the bias gradient of the Linear that fed x)
  int64_t x_stride, dy_stride, dx_stride;
  uint32_t rows, R;
  float eps;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v = __fadd_rn(v, __shfl_xor_sync(0xffffffffu, v, m));
  return v;
}
__device__ __forceinline__ float sum4(float4 a) { return __fadd_rn(__fadd_rn(a.x, a.y), __fadd_rn(a.z, a.w)); }
__device__ __forceinline__ float4 mul4(float4 a, float4 b) {
  return make_float4(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y), __fmul_rn(a.z, b.z), __fmul_rn(a.w, b.w));
}

__device__ __forceinline__ float block_sum(float v, float *scratch) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  const float r = (threadIdx.x & 31) < kWarps ? scratch[threadIdx.x & 31] : 0.f;
  return warp_sum(r);
}

// ------------------------------------------------------------------ softmax backward
template <int V>
__global__ void __launch_bounds__(kBlock) softmax_bwd_warp_kernel(const SoftmaxBwd P) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t r4 = P.R >> 2;
  for (uint32_t row = blockIdx.x * kWarps + warp; row < P.rows; row += gridDim.x * kWarps) {
    const float4 *yr = reinterpret_cast<const float4 *>(P.y + (int64_t)row * P.y_stride);
    const float4 *gr = reinterpret_cast<const float4 *>(P.dy + (int64_t)row * P.dy_stride);
    float4 *xr = reinterpret_cast<float4 *>(P.dx + (int64_t)row * P.dx_stride);
    float4 y[V], g[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const uint32_t i = lane + k * 32;
      y[k] = i < r4 ? __ldcs(yr + i) : make_float4(0, 0, 0, 0);
      g[k] = i < r4 ? __ldcs(gr + i) : make_float4(0, 0, 0, 0);
    }
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < V; ++k) dot = __fadd_rn(dot, sum4(mul4(g[k], y[k])));
    dot = warp_sum(dot);
    const uint32_t *mr = P.mask ? reinterpret_cast<const uint32_t *>(P.mask + (int64_t)(row % P.mask_rows) * P.mask_stride)
                                : nullptr;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const uint32_t i = lane + k * 32;
      if (i >= r4) continue;
      float4 o;
      o.x = __fmul_rn(__fsub_rn(g[k].x, dot), y[k].x); o.y = __fmul_rn(__fsub_rn(g[k].y, dot), y[k].y);
      o.z = __fmul_rn(__fsub_rn(g[k].z, dot), y[k].z); o.w = __fmul_rn(__fsub_rn(g[k].w, dot), y[k].w);
      if (P.mul != 0.0f) {
        o.x = __fmul_rn(o.x, P.mul); o.y = __fmul_rn(o.y, P.mul); o.z = __fmul_rn(o.z, P.mul); o.w = __fmul_rn(o.w, P.mul);
      } else if (P.div != 1.0f) {
        o.x = __fdiv_rn(o.x, P.div); o.y = __fdiv_rn(o.y, P.div); o.z = __fdiv_rn(o.z, P.div); o.w = __fdiv_rn(o.w, P.div);
      }
      if (mr) {
        const uint32_t m = __ldg(mr + i);
        if (m & 0xFFu) o.x = 0.f;
        if (m & 0xFF00u) o.y = 0.f;
        if (m & 0xFF0000u) o.z = 0.f;
        if (m & 0xFF000000u) o.w = 0.f;
      }
      __stcs(xr + i, o);
    }
  }
}

__global__ void __launch_bounds__(kBlock) softmax_bwd_cta_kernel(const SoftmaxBwd P) {
  __shared__ float scratch[kWarps];
  for (uint32_t row = blockIdx.x; row < P.rows; row += gridDim.x) {
    const float *yr = P.y + (int64_t)row * P.y_stride;
    const float *gr = P.dy + (int64_t)row * P.dy_stride;
    float *xr = P.dx + (int64_t)row * P.dx_stride;
    const uint8_t *mr = P.mask ? P.mask + (int64_t)(row % P.mask_rows) * P.mask_stride : nullptr;
    float dot = 0.f;
    for (uint32_t i = threadIdx.x; i < P.R; i += kBlock) dot = __fadd_rn(dot, __fmul_rn(gr[i], yr[i]));
    dot = block_sum(dot, scratch);
    for (uint32_t i = threadIdx.x; i < P.R; i += kBlock) {
      float o = __fmul_rn(__fsub_rn(gr[i], dot), yr[i]);
      if (P.div != 1.0f) o = __fdiv_rn(o, P.div);
      if (mr && mr[i]) o = 0.f;
      xr[i] = o;
    }
  }
}

// ------------------------------------------------------------------ layer_norm backward
template <int V>
__global__ void __launch_bounds__(kBlock, 2) layer_norm_bwd_warp_kernel(const LayerNormBwd P) {
  // [3][kWarps][r4]: every warp's running column sums of dy*xhat (dgamma), dy (dbeta) and dx, in SHARED memory.  As
  // registers (3 x V float4 per lane on top of the row's x and dy) they put the kernel at 191 registers and one 8-warp CTA
  // per SM — 12.6 % warps active, 20 % of the HBM rate (ncu); here two CTAs fit and the row loads of 16 warps overlap.
  extern __shared__ float4 part[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t r4 = P.R >> 2;
  float4 *sg = part + (size_t)warp * r4, *sb = part + (size_t)(kWarps + warp) * r4, *sd = part + (size_t)(2 * kWarps + warp) * r4;
  const bool want_dx = P.pdx != nullptr;
#pragma unroll
  for (int k = 0; k < V; ++k) {
    const uint32_t i = lane + k * 32;
    if (i < r4) sg[i] = sb[i] = sd[i] = make_float4(0, 0, 0, 0);     // a lane only ever touches its own slots: no sync
  }
  const float invR = 1.0f / (float)P.R;
  for (uint32_t row = blockIdx.x * kWarps + warp; row < P.rows; row += gridDim.x * kWarps) {
    const float4 *xr = reinterpret_cast<const float4 *>(P.x + (int64_t)row * P.x_stride);
    const float4 *gr = reinterpret_cast<const float4 *>(P.dy + (int64_t)row * P.dy_stride);
    float4 *dxr = reinterpret_cast<float4 *>(P.dx + (int64_t)row * P.dx_stride);
    float4 x[V], g[V];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const uint32_t i = lane + k * 32;
      x[k] = i < r4 ? __ldcs(xr + i) : make_float4(0, 0, 0, 0);
      g[k] = i < r4 ? __ldcs(gr + i) : make_float4(0, 0, 0, 0);
      s = __fadd_rn(s, sum4(x[k]));
    }
    const float mean = __fdiv_rn(warp_sum(s), (float)P.R);
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      x[k].x = __fsub_rn(x[k].x, mean); x[k].y = __fsub_rn(x[k].y, mean);
      x[k].z = __fsub_rn(x[k].z, mean); x[k].w = __fsub_rn(x[k].w, mean);
      if (lane + k * 32 < r4) q = __fadd_rn(q, sum4(mul4(x[k], x[k])));
    }
    const float var = __fdiv_rn(warp_sum(q), (float)P.R);
    const float denom = __fsqrt_rn(__fadd_rn(var, P.eps));
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const uint32_t i = lane + k * 32;
      if (i >= r4) continue;
      // xh, then the column partials with the raw dy, then g = dy*gamma
      x[k].x = __fdiv_rn(x[k].x, denom); x[k].y = __fdiv_rn(x[k].y, denom);
      x[k].z = __fdiv_rn(x[k].z, denom); x[k].w = __fdiv_rn(x[k].w, denom);
      const float4 dyx = mul4(g[k], x[k]);
      float4 ag = sg[i], ab = sb[i];
      ag.x = __fadd_rn(ag.x, dyx.x); ag.y = __fadd_rn(ag.y, dyx.y); ag.z = __fadd_rn(ag.z, dyx.z); ag.w = __fadd_rn(ag.w, dyx.w);
      ab.x = __fadd_rn(ab.x, g[k].x); ab.y = __fadd_rn(ab.y, g[k].y); ab.z = __fadd_rn(ab.z, g[k].z); ab.w = __fadd_rn(ab.w, g[k].w);
      sg[i] = ag;
      sb[i] = ab;
      if (P.gamma) g[k] = mul4(g[k], __ldg(reinterpret_cast<const float4 *>(P.gamma) + i));
      s1 = __fadd_rn(s1, sum4(g[k]));
      s2 = __fadd_rn(s2, sum4(mul4(g[k], x[k])));
    }
    const float m1 = __fmul_rn(warp_sum(s1), invR), m2 = __fmul_rn(warp_sum(s2), invR);
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const uint32_t i = lane + k * 32;
      if (i >= r4) continue;
      float4 o;
      o.x = __fdiv_rn(__fsub_rn(__fsub_rn(g[k].x, m1), __fmul_rn(x[k].x, m2)), denom);
      o.y = __fdiv_rn(__fsub_rn(__fsub_rn(g[k].y, m1), __fmul_rn(x[k].y, m2)), denom);
      o.z = __fdiv_rn(__fsub_rn(__fsub_rn(g[k].z, m1), __fmul_rn(x[k].z, m2)), denom);
      o.w = __fdiv_rn(__fsub_rn(__fsub_rn(g[k].w, m1), __fmul_rn(x[k].w, m2)), denom);
      __stcs(dxr + i, o);
      if (want_dx) {
        float4 ad = sd[i];
        ad.x = __fadd_rn(ad.x, o.x); ad.y = __fadd_rn(ad.y, o.y); ad.z = __fadd_rn(ad.z, o.z); ad.w = __fadd_rn(ad.w, o.w);
        sd[i] = ad;
      }
    }
  }
  // combine the CTA's warps (fixed order), one partial row per CTA
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < r4; i += kBlock) {
    float4 a = part[i], b2 = part[(size_t)kWarps * r4 + i], d = part[(size_t)2 * kWarps * r4 + i];
    for (int w = 1; w < kWarps; ++w) {
      const float4 a2 = part[(size_t)w * r4 + i], b3 = part[(size_t)(kWarps + w) * r4 + i], d2 = part[(size_t)(2 * kWarps + w) * r4 + i];
      a.x = __fadd_rn(a.x, a2.x); a.y = __fadd_rn(a.y, a2.y); a.z = __fadd_rn(a.z, a2.z); a.w = __fadd_rn(a.w, a2.w);
      b2.x = __fadd_rn(b2.x, b3.x); b2.y = __fadd_rn(b2.y, b3.y); b2.z = __fadd_rn(b2.z, b3.z); b2.w = __fadd_rn(b2.w, b3.w);
      d.x = __fadd_rn(d.x, d2.x); d.y = __fadd_rn(d.y, d2.y); d.z = __fadd_rn(d.z, d2.z); d.w = __fadd_rn(d.w, d2.w);
    }
    reinterpret_cast<float4 *>(P.pgamma + (size_t)blockIdx.x * P.R)[i] = a;
    reinterpret_cast<float4 *>(P.pbeta + (size_t)blockIdx.x * P.R)[i] = b2;
    if (want_dx) reinterpret_cast<float4 *>(P.pdx + (size_t)blockIdx.x * P.R)[i] = d;
  }
}

// generic R: one CTA per row, three streaming passes; column partials accumulated per thread
// (thread t owns columns t, t+kBlock, ...), so a CTA's partial row needs no cross-thread combine
__global__ void __launch_bounds__(kBlock) layer_norm_bwd_cta_kernel(const LayerNormBwd P) {
  __shared__ float scratch[kWarps];
  const float invR = 1.0f / (float)P.R;
  float *pg = P.pgamma + (size_t)blockIdx.x * P.R, *pb = P.pbeta + (size_t)blockIdx.x * P.R;
  float *pd = P.pdx ? P.pdx + (size_t)blockIdx.x * P.R : nullptr;
  for (uint32_t i = threadIdx.x; i < P.R; i += kBlock) {
    pg[i] = pb[i] = 0.f;
    if (pd) pd[i] = 0.f;
  }
  for (uint32_t row = blockIdx.x; row < P.rows; row += gridDim.x) {
    const float *xr = P.x + (int64_t)row * P.x_stride;
    const float *gr = P.dy + (int64_t)row * P.dy_stride;
    float *dxr = P.dx + (int64_t)row * P.dx_stride;
    float s = 0.f;
    for (uint32_t i = threadIdx.x; i < P.R; i += kBlock) s = __fadd_rn(s, xr[i]);
    const float mean = __fdiv_rn(block_sum(s, scratch), (float)P.R);
    float q = 0.f;
    for (uint32_t i = threadIdx.x; i < P.R; i += kBlock) {
      const float c = __fsub_rn(xr[i], mean);
      q = __fadd_rn(q, __fmul_rn(c, c));
    }
    const float var = __fdiv_rn(block_sum(q, scratch), (float)P.R);
    const float denom = __fsqrt_rn(__fadd_rn(var, P.eps));
    float s1 = 0.f, s2 = 0.f;
    for (uint32_t i = threadIdx.x; i < P.R; i += kBlock) {
      const float xh = __fdiv_rn(__fsub_rn(xr[i], mean), denom);
      const float dy = gr[i];
      pg[i] = __fadd_rn(pg[i], __fmul_rn(dy, xh));
      pb[i] = __fadd_rn(pb[i], dy);
      const float g = P.gamma ? __fmul_rn(dy, __ldg(P.gamma + i)) : dy;
      s1 = __fadd_rn(s1, g);
      s2 = __fadd_rn(s2, __fmul_rn(g, xh));
    }
    const float m1 = __fmul_rn(block_sum(s1, scratch), invR);
    const float m2 = __fmul_rn(block_sum(s2, scratch), invR);
    for (uint32_t i = threadIdx.x; i < P.R; i += kBlock) {
      const float xh = __fdiv_rn(__fsub_rn(xr[i], mean), denom);
      const float g = P.gamma ? __fmul_rn(gr[i], __ldg(P.gamma + i)) : gr[i];
      const float o = __fdiv_rn(__fsub_rn(__fsub_rn(g, m1), __fmul_rn(xh, m2)), denom);
      dxr[i] = o;
      if (pd) pd[i] = __fadd_rn(pd[i], o);
    }
  }
}

static int32_t rows_view(const b200_tensor &t, uint32_t &rows, uint32_t &R, int64_t &row_stride, const char *what) {
  B200_REQUIRE(t.rank >= 1 && t.rank <= B200_MAX_RANK, B200_ERR_INVALID, "%s: bad rank %d", what, t.rank);
  B200_REQUIRE(t.dtype == B200_F32 && t.ptr, B200_ERR_UNSUPPORTED, "%s: only f32 is implemented", what);
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

static bool aligned16(const void *p) { return ((uintptr_t)p) % 16 == 0; }

}  // namespace rnb
}  // namespace b200

using namespace b200;

extern "C" int32_t b200_launch_softmax_backward(const b200_tensor *y, const b200_tensor *dy, const b200_tensor *mask,
                                                double div, const b200_tensor *dx, b200_stream s) {
  B200_REQUIRE(y && dy && dx, B200_ERR_INVALID, "null argument");
  B200_REQUIRE(y->rank == dy->rank && y->rank == dx->rank, B200_ERR_SHAPE, "softmax_backward rank mismatch");
  for (int d = 0; d < y->rank; ++d)
    B200_REQUIRE(y->shape[d] == dy->shape[d] && y->shape[d] == dx->shape[d], B200_ERR_SHAPE,
                 "softmax_backward shape mismatch at dim %d", d);
  B200_REQUIRE(div != 0.0, B200_ERR_INVALID, "softmax_backward: div must be nonzero");
  rnb::SoftmaxBwd P;
  memset(&P, 0, sizeof(P));
  uint32_t rows2, R2;
  int32_t st = rnb::rows_view(*y, P.rows, P.R, P.y_stride, "softmax_backward y");
  if (st != B200_OK) return st;
  if ((st = rnb::rows_view(*dy, rows2, R2, P.dy_stride, "softmax_backward dy")) != B200_OK) return st;
  if ((st = rnb::rows_view(*dx, rows2, R2, P.dx_stride, "softmax_backward dx")) != B200_OK) return st;
  P.y = reinterpret_cast<const float *>(y->ptr);
  P.dy = reinterpret_cast<const float *>(dy->ptr);
  P.dx = reinterpret_cast<float *>(dx->ptr);
  P.div = (float)div;
  {
    int e = 0;
    const float m = frexpf(fabsf(P.div), &e);           // power of two with a normal reciprocal
    P.mul = (m == 0.5f && e > -120 && e < 120 && P.div != 1.0f) ? 1.0f / P.div : 0.0f;
  }
  bool mask_vec = true;
  if (mask) {
    // [..., Sq, R] with every leading dim of size 1 (broadcast over the batch), or the full shape
    B200_REQUIRE((mask->dtype == B200_BOOL || mask->dtype == B200_U8) && mask->ptr, B200_ERR_UNSUPPORTED, "mask must be bool");
    B200_REQUIRE(mask->rank == y->rank && mask->shape[mask->rank - 1] == y->shape[y->rank - 1] &&
                     (mask->shape[mask->rank - 1] == 1 || mask->strides[mask->rank - 1] == 1),
                 B200_ERR_SHAPE, "mask must match the last axis contiguously");
    int64_t mrows = 1, mstride = mask->shape[mask->rank - 1];
    bool first = true, full = true;
    for (int d = mask->rank - 2; d >= 0; --d) {
      if (mask->shape[d] == 1) { if (y->shape[d] != 1) full = false; continue; }
      B200_REQUIRE(mask->shape[d] == y->shape[d], B200_ERR_SHAPE, "mask dim %d is not broadcastable", d);
      B200_REQUIRE(full, B200_ERR_UNSUPPORTED, "mask may broadcast over leading dims only");
      if (first) { mstride = mask->strides[d]; first = false; }
      B200_REQUIRE(mask->strides[d] == mstride * mrows, B200_ERR_UNSUPPORTED, "mask dims must be jointly strided");
      mrows *= mask->shape[d];
    }
    P.mask = reinterpret_cast<const uint8_t *>(mask->ptr);
    P.mask_rows = (uint32_t)mrows;
    P.mask_stride = mstride;
    mask_vec = ((uintptr_t)mask->ptr) % 4 == 0 && mstride % 4 == 0;
  }
  if (P.rows == 0 || P.R == 0) return B200_OK;
  cudaStream_t stream = resolve_stream(s);
  const bool vec_ok = P.R % 4 == 0 && rnb::aligned16(P.y) && rnb::aligned16(P.dy) && rnb::aligned16(P.dx) &&
                      P.y_stride % 4 == 0 && P.dy_stride % 4 == 0 && P.dx_stride % 4 == 0 && mask_vec;
  const int sms = sm_count();
  if (vec_ok && P.R <= 2048) {
    const uint32_t r4 = P.R / 4;
    const unsigned grid = (unsigned)std::min<uint32_t>((P.rows + rnb::kWarps - 1) / rnb::kWarps, (uint32_t)sms * 8u);
    if (r4 <= 32) rnb::softmax_bwd_warp_kernel<1><<<grid, rnb::kBlock, 0, stream>>>(P);
    else if (r4 <= 64) rnb::softmax_bwd_warp_kernel<2><<<grid, rnb::kBlock, 0, stream>>>(P);
    else if (r4 <= 128) rnb::softmax_bwd_warp_kernel<4><<<grid, rnb::kBlock, 0, stream>>>(P);
    else if (r4 <= 256) rnb::softmax_bwd_warp_kernel<8><<<grid, rnb::kBlock, 0, stream>>>(P);
    else rnb::softmax_bwd_warp_kernel<16><<<grid, rnb::kBlock, 0, stream>>>(P);
  } else {
    const unsigned grid = (unsigned)std::min<uint32_t>(P.rows, (uint32_t)sms * 8u);
    rnb::softmax_bwd_cta_kernel<<<grid, rnb::kBlock, 0, stream>>>(P);
  }
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int32_t b200_layer_norm_backward_partials(const b200_tensor *input, int32_t *n_partials) {
  B200_REQUIRE(input && n_partials, B200_ERR_INVALID, "null argument");
  uint32_t rows, R;
  int64_t stride;
  int32_t st = rnb::rows_view(*input, rows, R, stride, "layer_norm_backward input");
  if (st != B200_OK) return st;
  const uint32_t per = R % 4 == 0 && R <= 2048 ? (uint32_t)rnb::kWarps : 1u;
  *n_partials = (int32_t)std::max<uint32_t>(1u, std::min<uint32_t>((rows + per - 1) / per, (uint32_t)sm_count() * 2u));
  return B200_OK;
}

extern "C" int32_t b200_launch_layer_norm_backward(const b200_tensor *input, const b200_tensor *dy,
                                                   const b200_tensor *gamma, double eps, const b200_tensor *dx,
                                                   const b200_tensor *partial_gamma, const b200_tensor *partial_beta,
                                                   b200_stream s) {
  return b200_launch_layer_norm_backward_ex(input, dy, gamma, eps, dx, partial_gamma, partial_beta, nullptr, s);
}

extern "C" int32_t b200_launch_layer_norm_backward_ex(const b200_tensor *input, const b200_tensor *dy,
                                                      const b200_tensor *gamma, double eps, const b200_tensor *dx,
                                                      const b200_tensor *partial_gamma, const b200_tensor *partial_beta,
                                                      const b200_tensor *partial_dx, b200_stream s) {
  B200_REQUIRE(input && dy && dx && partial_gamma && partial_beta, B200_ERR_INVALID, "null argument");
  B200_REQUIRE(input->rank == dy->rank && input->rank == dx->rank, B200_ERR_SHAPE, "layer_norm_backward rank mismatch");
  for (int d = 0; d < input->rank; ++d)
    B200_REQUIRE(input->shape[d] == dy->shape[d] && input->shape[d] == dx->shape[d], B200_ERR_SHAPE,
                 "layer_norm_backward shape mismatch at dim %d", d);
  rnb::LayerNormBwd P;
  memset(&P, 0, sizeof(P));
  uint32_t rows2, R2;
  int32_t st = rnb::rows_view(*input, P.rows, P.R, P.x_stride, "layer_norm_backward input");
  if (st != B200_OK) return st;
  if ((st = rnb::rows_view(*dy, rows2, R2, P.dy_stride, "layer_norm_backward dy")) != B200_OK) return st;
  if ((st = rnb::rows_view(*dx, rows2, R2, P.dx_stride, "layer_norm_backward dx")) != B200_OK) return st;
  int32_t G = 0;
  if ((st = b200_layer_norm_backward_partials(input, &G)) != B200_OK) return st;
  for (const b200_tensor *t : {partial_gamma, partial_beta, partial_dx}) {
    if (!t) continue;
    B200_REQUIRE(t->dtype == B200_F32 && t->ptr && t->rank == 2 && t->shape[0] == G && t->shape[1] == (int64_t)P.R &&
                     t->strides[1] == 1 && t->strides[0] == (int64_t)P.R,
                 B200_ERR_SHAPE, "layer_norm_backward partials must be contiguous f32 [%d, %u]", G, P.R);
  }
  if (gamma) {
    int64_t n = 1;
    for (int d = 0; d < gamma->rank; ++d) n *= gamma->shape[d];
    B200_REQUIRE(gamma->dtype == B200_F32 && gamma->ptr && n == (int64_t)P.R &&
                     (gamma->shape[gamma->rank - 1] == 1 || gamma->strides[gamma->rank - 1] == 1),
                 B200_ERR_SHAPE, "gamma must hold d_model = %u contiguous f32 values", P.R);
    P.gamma = reinterpret_cast<const float *>(gamma->ptr);
  }
  P.x = reinterpret_cast<const float *>(input->ptr);
  P.dy = reinterpret_cast<const float *>(dy->ptr);
  P.dx = reinterpret_cast<float *>(dx->ptr);
  P.pgamma = reinterpret_cast<float *>(partial_gamma->ptr);
  P.pbeta = reinterpret_cast<float *>(partial_beta->ptr);
  P.pdx = partial_dx ? reinterpret_cast<float *>(partial_dx->ptr) : nullptr;
  P.eps = (float)eps;
  cudaStream_t stream = resolve_stream(s);
  if (P.rows == 0 || P.R == 0) {
    B200_CUDA(cudaMemsetAsync(P.pgamma, 0, (size_t)G * P.R * 4, stream));
    B200_CUDA(cudaMemsetAsync(P.pbeta, 0, (size_t)G * P.R * 4, stream));
    if (P.pdx) B200_CUDA(cudaMemsetAsync(P.pdx, 0, (size_t)G * P.R * 4, stream));
    return B200_OK;
  }
  const bool vec_ok = P.R % 4 == 0 && P.R <= 2048 && rnb::aligned16(P.x) && rnb::aligned16(P.dy) && rnb::aligned16(P.dx) &&
                      rnb::aligned16(P.pgamma) && rnb::aligned16(P.pbeta) && (!P.pdx || rnb::aligned16(P.pdx)) && (!P.gamma || rnb::aligned16(P.gamma)) &&
                      P.x_stride % 4 == 0 && P.dy_stride % 4 == 0 && P.dx_stride % 4 == 0;
  if (vec_ok) {
    const uint32_t r4 = P.R / 4;
    const size_t smem = (size_t)3 * rnb::kWarps * r4 * sizeof(float4);
#define B200_LNB(Vv)                                                                                          \
  do {                                                                                                        \
    auto kern = rnb::layer_norm_bwd_warp_kernel<Vv>;                                                          \
    if (smem > 48 * 1024) B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<(unsigned)G, rnb::kBlock, smem, stream>>>(P);                                                      \
  } while (0)
    if (r4 <= 32) B200_LNB(1);
    else if (r4 <= 64) B200_LNB(2);
    else if (r4 <= 128) B200_LNB(4);
    else if (r4 <= 256) B200_LNB(8);
    else B200_LNB(16);
#undef B200_LNB
  } else {
    // the partial count was sized for the vector kernel only when R qualified; recompute for safety
    rnb::layer_norm_bwd_cta_kernel<<<(unsigned)G, rnb::kBlock, 0, stream>>>(P);
  }
  B200_LAUNCH_CHECK();
  return B200_OK;
}
