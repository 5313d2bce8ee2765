// Fast paths of path (b): plain f32 reductions of a contiguous tensor with no
// fused tape — float_sum / float_sum_dim / float_mean_dim / float_max_dim /
// float_argmax on materialised tensors, by far the most common calls
// (crates/burn-cubecl/src/kernel/reduce/base.rs:108-150).  Everything is
// specialised at compile time (kind, mapping), loads are 128-bit with four
// independent requests in flight per thread, and the combine tree is
// deterministic.  Semantics are those of reduce.cu (keepdim, first-index ties,
// first-NaN-wins, NaN-propagating max/min).
//
// Roofline: HBM; algorithmic bytes = input bytes + output bytes.
#pragma once
#include <cooperative_groups.h>

#include "tape.cuh"

namespace b200 {
namespace fast {

namespace cgf = cooperative_groups;

constexpr int kBlock = 256;
constexpr int kWarpsPerBlock = kBlock / 32;

enum Kind : int { kSum = 0, kMax = 1, kMin = 2, kArgMax = 3, kArgMin = 4 };

struct VI {
  float v;
  int32_t i;
};

template <int K>
__device__ __forceinline__ VI identity() {
  VI a;
  a.i = 0x7fffffff;
  a.v = (K == kSum) ? 0.0f : ((K == kMax || K == kArgMax) ? -INFINITY : INFINITY);
  return a;
}

// Folds element (x, idx) into a (sequential use: idx increases, but the rule is
// written to be order-independent so it also serves the tree combine).
template <int K>
__device__ __forceinline__ VI combine(VI a, VI b) {
  if constexpr (K == kSum) {
    a.v = __fadd_rn(a.v, b.v);
    return a;
  } else if constexpr (K == kMax || K == kMin) {
    const bool an = a.v != a.v, bn = b.v != b.v;
    float r = (K == kMax) ? fmaxf(a.v, b.v) : fminf(a.v, b.v);
    r = an ? a.v : (bn ? b.v : r);
    a.v = r;
    return a;
  } else {
    const bool an = a.v != a.v, bn = b.v != b.v;
    bool better = (K == kArgMax) ? (b.v > a.v) : (b.v < a.v);
    better = better || (b.v == a.v && b.i < a.i);
    const bool take_b = (an || bn) ? (bn && (!an || b.i < a.i)) : better;
    return take_b ? b : a;
  }
}

// Sequential fold: elements reach a thread in increasing index order, so a strict
// comparison keeps the first extreme, and "!(x <= best) && best == best" is both
// "x beats best" and "the first NaN wins and sticks" in two predicate instructions.
// Callers seed a.i with the index of the thread's first element (value = identity),
// so an all -inf (+inf) lane still reports its first index.
template <int K>
__device__ __forceinline__ VI fold(VI a, float x, int32_t idx) {
  if constexpr (K == kSum) {
    a.v = __fadd_rn(a.v, x);
  } else {
    const bool beats = (K == kMax || K == kArgMax) ? !(x <= a.v) : !(x >= a.v);
    const bool take = beats && (a.v == a.v);
    a.v = take ? x : a.v;
    if constexpr (K >= kArgMax) a.i = take ? idx : a.i;
  }
  return a;
}

// NaN-propagating extremes in one instruction (max.NaN / min.NaN, sm_80+)
__device__ __forceinline__ float max_nan(float a, float b) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float min_nan(float a, float b) {
  float r;
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}

// Folds a batch of four vectors (element index of v[k].x = first[k], increasing in k; ok[k] = vector in range; vectors out
// of range hold the identity).  Sum: 16 adds.  Extremes: ONE instruction per element — the batch extreme by a
// NaN-propagating tree — and one test per batch: only when the batch holds something that beats the running value
// (a new extreme, or the first NaN) is it walked element by element with the sequential rule, which is what keeps
// "first extreme / first NaN wins" exact.  For random data a thread takes the slow walk O(log n) times.
template <int K>
__device__ __forceinline__ VI fold_batch(VI a, const float4 (&v)[4], const int32_t (&first)[4], const bool (&ok)[4]) {
  if constexpr (K == kSum) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (ok[k]) {
        a.v = __fadd_rn(a.v, v[k].x); a.v = __fadd_rn(a.v, v[k].y); a.v = __fadd_rn(a.v, v[k].z); a.v = __fadd_rn(a.v, v[k].w);
      }
    return a;
  } else {
    constexpr bool kIsMax = (K == kMax || K == kArgMax);
    auto ext = [](float x, float y) { return kIsMax ? max_nan(x, y) : min_nan(x, y); };
    float m[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) m[k] = ext(ext(v[k].x, v[k].y), ext(v[k].z, v[k].w));
    const float mm = ext(ext(m[0], m[1]), ext(m[2], m[3]));
    if constexpr (K == kMax || K == kMin) {
      a.v = ext(a.v, mm);                       // NaN sticks: ext(NaN, x) = NaN
      return a;
    } else {
      const bool beats = kIsMax ? !(mm <= a.v) : !(mm >= a.v);
      if (beats && a.v == a.v) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (ok[k]) {
            a = fold<K>(a, v[k].x, first[k]);
            a = fold<K>(a, v[k].y, first[k] + 1);
            a = fold<K>(a, v[k].z, first[k] + 2);
            a = fold<K>(a, v[k].w, first[k] + 3);
          }
      }
      return a;
    }
  }
}

// (value, index) packed into one 64-bit key whose unsigned maximum IS the arg-reduction rule: the high word orders
// the values (monotone image of the float, complemented for argmin, NaN above everything, -0 folded onto +0 because
// they compare equal), the low word is ~index so that among equal values the smallest index wins.  The tree combine
// of an arg reduction is then two shuffles and one 64-bit max per step instead of a dozen predicated instructions.
template <int K>
__device__ __forceinline__ unsigned long long vi_key(VI a) {
  uint32_t u = __float_as_uint(a.v);
  u = (a.v == 0.0f) ? 0u : u;
  uint32_t k = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  if (K == kArgMin) k = ~k;
  k = (a.v != a.v) ? 0xFFFFFFFFu : k;
  return ((unsigned long long)k << 32) | (unsigned long long)(~(uint32_t)a.i);
}
template <int K>
__device__ __forceinline__ VI vi_unkey(unsigned long long key) {
  uint32_t k = (uint32_t)(key >> 32);
  VI a;
  a.i = (int32_t)(~(uint32_t)key);
  if (k == 0xFFFFFFFFu) {
    a.v = __int_as_float(0x7fc00000);
  } else {
    if (K == kArgMin) k = ~k;
    a.v = __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
  }
  return a;
}

template <int K>
__device__ __forceinline__ VI warp_reduce(VI a) {
  if constexpr (K >= kArgMax) {
    unsigned long long key = vi_key<K>(a);
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, m);
      key = other > key ? other : key;
    }
    return vi_unkey<K>(key);
  } else {
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
      VI b;
      b.v = __shfl_xor_sync(0xffffffffu, a.v, m);
      b.i = 0;
      a = combine<K>(a, b);
    }
    return a;
  }
}

template <int K>
__device__ __forceinline__ VI block_reduce(VI a, VI *scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  a = warp_reduce<K>(a);
  __syncthreads();
  if (lane == 0) scratch[warp] = a;
  __syncthreads();
  VI r = (lane < kWarpsPerBlock) ? scratch[lane] : identity<K>();
  return warp_reduce<K>(r);
}

template <int K>
__device__ __forceinline__ void store_result(void *out, int32_t out_dtype, int64_t idx, VI a, bool mean, float div) {
  if constexpr (K >= kArgMax) {
    if (out_dtype == B200_I64) reinterpret_cast<long long *>(out)[idx] = (long long)a.i;
    else reinterpret_cast<int32_t *>(out)[idx] = a.i;
  } else {
    float v = a.v;
    if (mean) v = __fdiv_rn(v, div);
    store_one(out, out_dtype, idx, u_of(v));
  }
}

struct RowParams {
  const float *x;
  void *out;
  int32_t out_dtype;
  uint32_t n_rows;
  uint32_t r4;          // row length in float4
  uint32_t splits;      // CTAs per row
  uint32_t per_split;   // float4 per split (multiple of kBlock*4)
  VI *partials;
  uint32_t *tickets;
  int32_t mean;
  float div;
};

// One CTA per (row, split); four 128-bit loads in flight per thread.
template <int K>
__global__ void __launch_bounds__(kBlock) reduce_row_fast_kernel(const RowParams P) {
  __shared__ VI scratch[kWarpsPerBlock];
  __shared__ uint32_t s_last;
  const uint32_t n_work = P.n_rows * P.splits;
  for (uint32_t w = blockIdx.x; w < n_work; w += gridDim.x) {
    const uint32_t row = w / P.splits, split = w - row * P.splits;
    const float4 *p = reinterpret_cast<const float4 *>(P.x) + (size_t)row * P.r4;
    const uint32_t begin = split * P.per_split, end = min(P.r4, begin + P.per_split);
    VI a = identity<K>();
    if (begin + threadIdx.x < end) a.i = (int32_t)((begin + threadIdx.x) * 4);
    for (uint32_t base = begin + threadIdx.x; base < end; base += kBlock * 4) {
      float4 v[4];
      int32_t first[4];
      bool ok[4];
      const float idv = identity<K>().v;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t i = base + k * kBlock;
        ok[k] = i < end;
        first[k] = (int32_t)(i * 4);
        v[k] = ok[k] ? __ldcs(p + i) : make_float4(idv, idv, idv, idv);
      }
      a = fold_batch<K>(a, v, first, ok);
    }
    a = block_reduce<K>(a, scratch);
    if (P.splits == 1) {
      if (threadIdx.x == 0) store_result<K>(P.out, P.out_dtype, row, a, P.mean, P.div);
    } else {
      if (threadIdx.x == 0) {
        P.partials[(size_t)row * P.splits + split] = a;
        __threadfence();
        s_last = (atomicAdd(&P.tickets[row], 1u) == P.splits - 1) ? 1u : 0u;
      }
      __syncthreads();
      if (s_last) {
        __threadfence();
        VI b = identity<K>();
        for (uint32_t s = threadIdx.x; s < P.splits; s += kBlock) {
          const float2 raw = __ldcg(reinterpret_cast<const float2 *>(&P.partials[(size_t)row * P.splits + s]));
          VI e;
          e.v = raw.x;
          e.i = __float_as_int(raw.y);
          b = combine<K>(b, e);
        }
        b = block_reduce<K>(b, scratch);
        if (threadIdx.x == 0) {
          store_result<K>(P.out, P.out_dtype, row, b, P.mean, P.div);
          P.tickets[row] = 0u;
        }
      }
    }
    __syncthreads();
  }
}

// One warp per short row (R <= 4096 elements): no shared memory, no barriers.
template <int K>
__global__ void __launch_bounds__(kBlock) reduce_row_warp_fast_kernel(const RowParams P) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t row = blockIdx.x * kWarpsPerBlock + warp; row < P.n_rows; row += gridDim.x * kWarpsPerBlock) {
    const float4 *p = reinterpret_cast<const float4 *>(P.x) + (size_t)row * P.r4;
    VI a = identity<K>();
    if ((uint32_t)lane < P.r4) a.i = lane * 4;
    for (uint32_t base = lane; base < P.r4; base += 32 * 4) {
      float4 v[4];
      int32_t first[4];
      bool ok[4];
      const float idv = identity<K>().v;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t i = base + k * 32;
        ok[k] = i < P.r4;
        first[k] = (int32_t)(i * 4);
        v[k] = ok[k] ? __ldcs(p + i) : make_float4(idv, idv, idv, idv);
      }
      a = fold_batch<K>(a, v, first, ok);
    }
    a = warp_reduce<K>(a);
    if (lane == 0) store_result<K>(P.out, P.out_dtype, row, a, P.mean, P.div);
  }
}

struct ColParams {
  const float *x;
  void *out;
  int32_t out_dtype;
  uint32_t outer, R, inner4;   // inner in float4
  uint32_t splits, rows_per_split;
  int32_t mean;
  float div;
  uint32_t tx;                 // column-vectors per CTA (32, 16 or 8)
  float *partials;             // sum kinds, tall narrow inputs: [outer, splits, inner] partial sums instead of a cluster combine
};

// blockDim = (32, 8): 32 column-vectors (128 columns, 512 B per row) x 8 row
// groups; gridDim = (column tiles, splits) with cluster (1, splits, 1) — the
// cluster's partial columns are combined through distributed shared memory.
// TX narrows to 16 / 8 column-vectors (TY = 16 / 32 row groups) when 128-column tiles would
// leave most SMs without a CTA (tall, narrow inputs such as a [8192, 1024] bias gradient).
template <int K, int TX>
__global__ void __launch_bounds__(kBlock) reduce_col_fast_kernel(const ColParams P) {
  constexpr int TY = kBlock / TX;
  __shared__ VI part[TY][TX][4];
  __shared__ VI cta_result[TX][4];
  const uint32_t tiles_per_outer = (P.inner4 + TX - 1) / TX;
  const uint32_t o = blockIdx.x / tiles_per_outer;
  const uint32_t c4 = (blockIdx.x - o * tiles_per_outer) * TX + threadIdx.x;
  const bool col_ok = c4 < P.inner4;
  const uint32_t r_begin = blockIdx.y * P.rows_per_split;
  const uint32_t r_end = min(P.R, r_begin + P.rows_per_split);
  const float4 *p = reinterpret_cast<const float4 *>(P.x) + (size_t)o * P.R * P.inner4 + c4;

  VI a[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    a[j] = identity<K>();
    if (r_begin + threadIdx.y < r_end) a[j].i = (int32_t)(r_begin + threadIdx.y);
  }
  if (col_ok) {
    for (uint32_t base = r_begin + threadIdx.y; base < r_end; base += TY * 4) {
      float4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t r = base + k * TY;
        v[k] = r < r_end ? __ldcs(p + (size_t)r * P.inner4) : make_float4(0, 0, 0, 0);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t r = base + k * TY;
        if (r < r_end) {
          a[0] = fold<K>(a[0], v[k].x, (int32_t)r);
          a[1] = fold<K>(a[1], v[k].y, (int32_t)r);
          a[2] = fold<K>(a[2], v[k].z, (int32_t)r);
          a[3] = fold<K>(a[3], v[k].w, (int32_t)r);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) part[threadIdx.y][threadIdx.x][j] = a[j];
  __syncthreads();
  if (threadIdx.y == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      VI b = part[0][threadIdx.x][j];
      for (int y = 1; y < TY; ++y) b = combine<K>(b, part[y][threadIdx.x][j]);
      a[j] = b;
      cta_result[threadIdx.x][j] = b;
    }
  }
  if (P.partials) {
    // Tall, narrow inputs ([16384, 512] bias gradients): a cluster holds at most 8 row splits, i.e. 128 CTAs for 16 column
    // tiles — under one CTA per SM and 2 TB/s.  Up to 64 splits write partial rows here and the short-axis kernel
    // (reduce_col_short_kernel, fixed order) finishes them.
    if (threadIdx.y == 0 && col_ok) {
      float4 r = make_float4(a[0].v, a[1].v, a[2].v, a[3].v);
      reinterpret_cast<float4 *>(P.partials)[((size_t)o * P.splits + blockIdx.y) * P.inner4 + c4] = r;
    }
    return;
  }
  if (P.splits > 1) {
    cgf::cluster_group cluster = cgf::this_cluster();
    cluster.sync();
    if (cluster.block_rank() == 0 && threadIdx.y == 0) {
      for (uint32_t rk = 1; rk < P.splits; ++rk) {
        const VI *remote = cluster.map_shared_rank(&cta_result[0][0], rk);
#pragma unroll
        for (int j = 0; j < 4; ++j) a[j] = combine<K>(a[j], remote[threadIdx.x * 4 + j]);
      }
    }
    cluster.sync();
    if (cluster.block_rank() != 0) return;
  }
  if (threadIdx.y == 0 && col_ok) {
    const int64_t base = ((int64_t)o * P.inner4 + c4) * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) store_result<K>(P.out, P.out_dtype, base + j, a[j], P.mean, P.div);
  }
}

// Short reduced axis over many columns ([outer, R <= 64, inner] with inner in the hundreds of thousands: the split-K
// combine of a weight-gradient GEMM, per-CTA partial rows of a fused kernel): one thread per column-vector walks all R
// rows with four loads in flight — no shared memory, no cluster; the tiled kernel above would give every CTA 4 KB.
template <int K>
__global__ void __launch_bounds__(kBlock) reduce_col_short_kernel(const ColParams P) {
  const uint64_t total = (uint64_t)P.outer * P.inner4;
  for (uint64_t g = (uint64_t)blockIdx.x * kBlock + threadIdx.x; g < total; g += (uint64_t)gridDim.x * kBlock) {
    const uint32_t o = (uint32_t)(g / P.inner4), c4 = (uint32_t)(g - (uint64_t)o * P.inner4);
    const float4 *p = reinterpret_cast<const float4 *>(P.x) + (size_t)o * P.R * P.inner4 + c4;
    VI a[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      a[j] = identity<K>();
      a[j].i = 0;
    }
    for (uint32_t base = 0; base < P.R; base += 4) {
      float4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = base + k < P.R ? __ldcs(p + (size_t)(base + k) * P.inner4) : make_float4(0, 0, 0, 0);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (base + k < P.R) {
          a[0] = fold<K>(a[0], v[k].x, (int32_t)(base + k));
          a[1] = fold<K>(a[1], v[k].y, (int32_t)(base + k));
          a[2] = fold<K>(a[2], v[k].z, (int32_t)(base + k));
          a[3] = fold<K>(a[3], v[k].w, (int32_t)(base + k));
        }
    }
    const int64_t ob = ((int64_t)o * P.inner4 + c4) * 4;
#pragma unroll
    for (int j = 0; j < 4; ++j) store_result<K>(P.out, P.out_dtype, ob + j, a[j], P.mean, P.div);
  }
}

}  // namespace fast
}  // namespace b200
