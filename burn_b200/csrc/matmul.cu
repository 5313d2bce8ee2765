// Path (c): float_matmul as a TMA-fed tcgen05 GEMM with TMEM accumulators and a
// fused fuse-on-write epilogue tape.
//
// Replaces matmul / launch_matmul → cubek::matmul (crates/burn-cubecl/src/kernel/matmul/base.rs:53-140)
// and MatmulOptimization::execute (crates/burn-cubecl-fusion/src/optim/matmul/optimization.rs:97-140).
// Semantics: C[..., M, N] = A[..., M, K] · B[..., K, N], numpy-style broadcast of the leading
// dims, operands may be transposed views (crates/burn-ndarray/src/ops/matmul.rs:9-183).
//
// Kernel (one CTA per SM, persistent over 128x128 output tiles, 192 threads):
//   warp 0   TMA producer: cp.async.bulk.tensor tiles of A and B into a K-stage smem ring
//            (SWIZZLE_128B; K-major or MN-major operands are both loaded in their native
//            memory orientation, so NN / NT / TN / TT need no transposing copy)
//   warp 1   MMA issuer: one elected thread issues tcgen05.mma (kind::tf32 or kind::f16 with
//            bf16 inputs, f32 accumulate) into one of two 128-column TMEM accumulators and
//            commits to mbarriers; also owns TMEM alloc/dealloc
//   warps 2-5 epilogue: tcgen05.ld the finished accumulator (lane = output row), run the
//            epilogue tape in registers (bias add, gelu chain, …) and store; overlaps the next
//            tile's MMAs thanks to the second accumulator
// Precisions: TF32 (f32 operands read directly, mantissa truncated by the tensor core),
// BF16 (operands rounded to bf16 in a pre-pass), F32X3 (3xTF32: hi/lo split operands
// concatenated along K so A_hi·B_hi + A_hi·B_lo + A_lo·B_hi accumulates in one pass).
// Roofline: tensor pipe (dense bf16 / tf32 peak); algorithmic FLOPs = 2·M·N·K per batch.
#include <cuda.h>

#include "tape_host.cuh"

namespace b200 {
namespace mm {

constexpr int BM = 128, BN = 128;
constexpr int kThreads = 192;
constexpr int kAccStages = 2;
constexpr int kMaxBatchDims = 3;
constexpr int kEpiBlock = 128, kEpiU = 4;  // epilogue tape geometry: 16 columns per dispatch

struct Params {
  CUtensorMap tma_a, tma_b;
  float *c;
  int64_t c_batch_stride[kMaxBatchDims];
  int64_t ldc;
  int32_t M, N, K;
  int32_t batch[kMaxBatchDims];   // collapsed output batch dims (1 when unused)
  int32_t a_bflag[kMaxBatchDims], b_bflag[kMaxBatchDims];  // 0 = operand broadcast along the dim
  int32_t tiles_m, tiles_n, k_blocks, stages;
  int32_t has_epilogue;
  int32_t c_dtype;
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_5d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1,
                                            int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
template <bool BF16>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  if constexpr (BF16) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc)
        : "memory");
  } else {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc)
        : "memory");
  }
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(pred));
  return pred != 0;
}

// Shared-memory matrix descriptor (sm_100 format, version 1).  layout: 2 = SWIZZLE_128B (16-byte
// swizzle atoms), 1 = SWIZZLE_128B_BASE32B (32-byte atoms — what MN-major 32-bit operands need;
// the TMA side is CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)layout << 61;
  return d;
}

// ------------------------------------------------------------------ kernel
// ES = operand element size (4: tf32, 2: bf16).  A_MN / B_MN: operand is MN-major in memory.
template <int ES, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ Params P, const __grid_constant__ TapeParams T) {
  constexpr int BK = 128 / ES;            // K elements per stage (one 128-byte swizzle row)
  constexpr int UMMA_K = 32 / ES;         // K per tcgen05.mma
  constexpr uint32_t kATile = BM * BK * ES, kBTile = BN * BK * ES;  // 16 KB each
  constexpr uint32_t kStageBytes = kATile + kBTile;
  constexpr bool BF16 = ES == 2;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *tiles = smem;
  uint64_t *full = reinterpret_cast<uint64_t *>(tiles + (size_t)P.stages * kStageBytes);
  uint64_t *empty = full + P.stages;
  uint64_t *acc_full = empty + P.stages;
  uint64_t *acc_empty = acc_full + kAccStages;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + kAccStages);
  uint32_t *epi_smem = reinterpret_cast<uint32_t *>(
      (reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_batch = P.batch[0] * P.batch[1] * P.batch[2];
  const int n_tiles = P.tiles_m * P.tiles_n * n_batch;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tma_a);
    tma_prefetch_desc(&P.tma_b);
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, kAccStages * BN);
  if (P.has_epilogue && warp >= 2) {
    SlotFile<4, kEpiU, kEpiBlock> slots;
    slots.smem = epi_smem;
    slots.tid = threadIdx.x - 64;
    init_scalars<4, kEpiU, kEpiBlock>(T, slots, T.n_in + T.n_tmp);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto tile_coords = [&](int tile, int &m_blk, int &n_blk, int (&b)[kMaxBatchDims]) {
    m_blk = tile % P.tiles_m;
    int r = tile / P.tiles_m;
    n_blk = r % P.tiles_n;
    r /= P.tiles_n;
    b[2] = r % P.batch[2];
    r /= P.batch[2];
    b[1] = r % P.batch[1];
    b[0] = r / P.batch[1];
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        int m_blk, n_blk, b[kMaxBatchDims];
        tile_coords(tile, m_blk, n_blk, b);
        const int ab0 = b[0] * P.a_bflag[0], ab1 = b[1] * P.a_bflag[1], ab2 = b[2] * P.a_bflag[2];
        const int bb0 = b[0] * P.b_bflag[0], bb1 = b[1] * P.b_bflag[1], bb2 = b[2] * P.b_bflag[2];
        for (int kb = 0; kb < P.k_blocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t *sa = tiles + (size_t)stage * kStageBytes;
          uint8_t *sb = sa + kATile;
          mbar_expect_tx(&full[stage], kStageBytes);
          if constexpr (!A_MN) {
            tma_load_5d(sa, &P.tma_a, &full[stage], kb * BK, m_blk * BM, ab2, ab1, ab0);
          } else {
#pragma unroll
            for (int g = 0; g < BM / BK; ++g)  // BK == elements per 128-byte MN group
              tma_load_5d(sa + g * (BK * 128), &P.tma_a, &full[stage], m_blk * BM + g * BK, kb * BK, ab2, ab1, ab0);
          }
          if constexpr (!B_MN) {
            tma_load_5d(sb, &P.tma_b, &full[stage], kb * BK, n_blk * BN, bb2, bb1, bb0);
          } else {
#pragma unroll
            for (int g = 0; g < BN / BK; ++g)
              tma_load_5d(sb + g * (BK * 128), &P.tma_b, &full[stage], n_blk * BN + g * BK, kb * BK, bb2, bb1, bb0);
          }
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // instruction descriptor: D = F32, A/B = TF32 or BF16, majors, N, M
    uint32_t idesc = 0;
    idesc |= 1u << 4;                              // c_format = F32
    idesc |= (BF16 ? 1u : 2u) << 7;                // a_format
    idesc |= (BF16 ? 1u : 2u) << 10;               // b_format
    idesc |= (A_MN ? 1u : 0u) << 15;               // a_major
    idesc |= (B_MN ? 1u : 0u) << 16;               // b_major
    idesc |= (uint32_t)(BN >> 3) << 17;            // n_dim
    idesc |= (uint32_t)(BM >> 4) << 24;            // m_dim
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      mbar_wait(&acc_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
      for (int kb = 0; kb < P.k_blocks; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(tiles + (size_t)stage * kStageBytes);
          const uint32_t sb = sa + kATile;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // K-major: 32 bytes along the swizzled row per MMA; MN-major: UMMA_K rows of 128 B
            // MN-major (validated on hardware for bf16): LBO = stride between 128-byte MN groups
            // (one TMA box of BK rows), SBO = stride between 8-row K atoms
            // (one TMA box of BK rows), SBO = stride between K atoms: 8 rows of 128 B for 16-bit
            // operands, 4 rows for 32-bit ones (32-byte swizzle atoms)
            constexpr uint32_t kMnSbo = BF16 ? 1024 : 512, kMnLayout = BF16 ? 2 : 1;
            const uint64_t da = A_MN ? make_desc(sa + k * (UMMA_K * 128), BK * 128, kMnSbo, kMnLayout)
                                     : make_desc(sa + k * 32, 16, 1024);
            const uint64_t db = B_MN ? make_desc(sb + k * (UMMA_K * 128), BK * 128, kMnSbo, kMnLayout)
                                     : make_desc(sb + k * 32, 16, 1024);
            umma<BF16>(tmem_d, da, db, idesc, (kb | k) ? 1u : 0u);
          }
        }
        __syncwarp();
        if (lane == 0) umma_commit(&empty[stage]);  // frees the smem stage when these MMAs retire
        if (++stage == P.stages) { stage = 0; phase ^= 1; }
      }
      if (lane == 0) umma_commit(&acc_full[acc]);   // accumulator complete → epilogue
      __syncwarp();
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;                  // TMEM lane quarter this warp may access
    const int row_in_tile = quarter * 32 + lane;
    SlotFile<4, kEpiU, kEpiBlock> slots;
    slots.smem = epi_smem;
    slots.tid = threadIdx.x - 64;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      int m_blk, n_blk, b[kMaxBatchDims];
      tile_coords(tile, m_blk, n_blk, b);
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const int m = m_blk * BM + row_in_tile;
      const int64_t c_batch = b[0] * P.c_batch_stride[0] + b[1] * P.c_batch_stride[1] + b[2] * P.c_batch_stride[2];
      float *crow = P.c + c_batch + (int64_t)m * P.ldc;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN);
      const int batch_lin = (b[0] * P.batch[1] + b[1]) * P.batch[2] + b[2];
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + c0, r);
        tmem_ld_wait();
        const int n0 = n_blk * BN + c0;
        if (m >= P.M || n0 >= P.N) continue;
        if (!P.has_epilogue) {
          if (n0 + 16 <= P.N && (reinterpret_cast<uintptr_t>(crow + n0) & 15) == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *reinterpret_cast<uint4 *>(crow + n0 + q * 4) = make_uint4(r[q * 4], r[q * 4 + 1], r[q * 4 + 2], r[q * 4 + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (n0 + j < P.N) crow[n0 + j] = f_of(r[j]);
          }
        } else {
          // fuse-on-write tape: INPUT(0) = accumulator, INPUT(k) = epilogue operands at (batch, m, n)
          uint32_t acc_v[kEpiU][4];
#pragma unroll
          for (int u = 0; u < kEpiU; ++u) {
#pragma unroll
            for (int j = 0; j < 4; ++j) acc_v[u][j] = r[u * 4 + j];
            slots.put(0, u, acc_v[u]);
            Coord3 c;
            c.c0 = (uint32_t)batch_lin;
            c.c1 = (uint32_t)m;
            c.c2 = (uint32_t)min(n0 + u * 4, P.N - 4 < 0 ? 0 : P.N - 4);
            for (int k = 1; k < T.n_in; ++k) {
              uint32_t v[4];
              load_operand<4, kRank3>(T, T.in[k], 0u, c, v);
              slots.put(k, u, v);
            }
          }
          run_tape<4, kEpiU, kEpiBlock>(
              T, slots, acc_v,
              [&](int o, const uint32_t(&val)[kEpiU][4]) {
                const OperandDesc &d = T.out[o];
#pragma unroll
                for (int u = 0; u < kEpiU; ++u) {
                  const int n = n0 + u * 4;
                  if (n + 4 <= P.N) {
                    Coord3 c;
                    c.c0 = (uint32_t)batch_lin;
                    c.c1 = (uint32_t)m;
                    c.c2 = (uint32_t)n;
                    store_operand<4, kRank3>(T, d, 0u, c, val[u]);
                  } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                      if (n + j < P.N) {
                        const int64_t off = (int64_t)batch_lin * d.s3[0] + (int64_t)m * d.s3[1] + (int64_t)(n + j) * d.s3[2];
                        store_one(d.ptr, d.dtype, off, val[u][j]);
                      }
                  }
                }
              },
              0, T.n_in);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kAccStages * BN);
  }
}

// ------------------------------------------------------------------ operand preparation kernels
// 3xTF32 split: writes [hi | hi | lo] (which = 0, operand A) or [hi | lo | hi] (which = 1,
// operand B) along K' = 3K as a K-major matrix [rows, 3K] from a strided [rows, K] view.
struct SplitBatch {
  int32_t bsz[kMaxBatchDims];
  int64_t s_b[kMaxBatchDims];
};
__global__ void split3_kernel(const float *src, int64_t s_row, int64_t s_k, SplitBatch sb, float *dst, int rows, int K,
                              int K3pad, int which, int64_t n_total) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const int64_t t = i / K;
    const int r = (int)(t % rows);
    const int64_t bt = t / rows;
    int64_t rest = bt, boff = 0;
#pragma unroll
    for (int d = kMaxBatchDims - 1; d >= 0; --d) {
      const int64_t c = rest % sb.bsz[d];
      rest /= sb.bsz[d];
      boff += c * sb.s_b[d];
    }
    const float x = src[boff + (int64_t)r * s_row + (int64_t)k * s_k];
    uint32_t hi_b, lo_b;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi_b) : "f"(x));
    const float hi = __uint_as_float(hi_b);
    const float rest_f = x - hi;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo_b) : "f"(rest_f));
    const float lo = __uint_as_float(lo_b);
    float *row = dst + (bt * rows + r) * (int64_t)K3pad;
    row[k] = hi;
    row[K + k] = which == 0 ? hi : lo;
    row[2 * K + k] = which == 0 ? lo : hi;
  }
}

}  // namespace mm

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// One GEMM operand as the kernel sees it: a [MN, K] matrix per batch element.
struct Operand {
  void *ptr;
  int es;              // element size
  int64_t s_mn, s_k;   // element strides
  int64_t s_b[mm::kMaxBatchDims];
  int32_t bsz[mm::kMaxBatchDims];  // operand's own batch extents (1 = broadcast)
  bool mn_major;
};

static bool tma_ok(const Operand &o, int64_t mn, int64_t k) {
  const int64_t align = 16 / o.es;
  // MN-major operands are consumed in place through MN-major UMMA descriptors: 16-bit ones with the
  // plain SWIZZLE_128B layout, 32-bit ones with SWIZZLE_128B_BASE32B (TMA: SWIZZLE_128B_ATOM_32B).
  // B200_MM_REPACK_TF32_MN=1 forces the old K-major repack pre-pass for 32-bit operands (debug).
  static const bool repack32 = std::getenv("B200_MM_REPACK_TF32_MN") != nullptr;
  if (o.mn_major && o.es == 4 && repack32) return false;
  const int64_t inner = o.mn_major ? o.s_mn : o.s_k, outer = o.mn_major ? o.s_k : o.s_mn;
  if (inner != 1) return false;
  if ((o.mn_major ? k : mn) > 1 && (outer % align != 0 || outer < (o.mn_major ? mn : k))) return false;
  for (int d = 0; d < mm::kMaxBatchDims; ++d)
    if (o.bsz[d] > 1 && (o.s_b[d] % align != 0 || o.s_b[d] <= 0)) return false;
  return true;
}

static int32_t make_tmap(CUtensorMap *map, const Operand &o, int64_t mn, int64_t k) {
  EncodeTiledFn enc = encode_fn();
  B200_REQUIRE(enc, B200_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable from the driver");
  const int BK = 128 / o.es;
  cuuint64_t dims[5], strides[4];
  cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
  const int64_t outer_stride = o.mn_major ? o.s_k : o.s_mn;
  dims[0] = (cuuint64_t)(o.mn_major ? mn : k);
  dims[1] = (cuuint64_t)(o.mn_major ? k : mn);
  box[0] = (cuuint32_t)BK;                       // 128 bytes of the contiguous dimension
  box[1] = (cuuint32_t)(o.mn_major ? BK : 128);  // MN-major: BK k-rows; K-major: 128 MN rows
  strides[0] = (cuuint64_t)std::max<int64_t>(outer_stride, 16 / o.es) * o.es;
  // tensor-map dims 2,3,4 = batch dims innermost-last → (b2, b1, b0)
  for (int d = 0; d < mm::kMaxBatchDims; ++d) {
    const int src = mm::kMaxBatchDims - 1 - d;
    dims[2 + d] = (cuuint64_t)std::max(o.bsz[src], 1);
    box[2 + d] = 1;
    const int64_t sb = o.bsz[src] > 1 ? o.s_b[src] : (int64_t)(16 / o.es);
    strides[1 + d] = (cuuint64_t)sb * o.es;
  }
  const CUtensorMapDataType dt = o.es == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const CUtensorMapSwizzle sw = (o.mn_major && o.es == 4) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = enc(map, dt, 5, o.ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_REQUIRE(r == CUDA_SUCCESS, B200_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
  return B200_OK;
}

struct MatmulPlan {
  int rank, nb;
  int64_t M, N, K;
  int32_t batch[mm::kMaxBatchDims];
  int64_t a_sb[mm::kMaxBatchDims], b_sb[mm::kMaxBatchDims], c_sb[mm::kMaxBatchDims];
  int32_t a_bsz[mm::kMaxBatchDims], b_bsz[mm::kMaxBatchDims];
};

static int32_t plan_matmul(const b200_tensor *a, const b200_tensor *b, const b200_tensor *c, MatmulPlan &pl) {
  B200_REQUIRE(a && b, B200_ERR_INVALID, "null operand");
  B200_REQUIRE(a->rank == b->rank && a->rank >= 2 && a->rank <= 2 + mm::kMaxBatchDims, B200_ERR_SHAPE,
               "matmul operands must have the same rank in [2, %d] (got %d and %d)", 2 + mm::kMaxBatchDims, a->rank,
               b->rank);
  const int r = a->rank;
  pl.rank = r;
  pl.nb = r - 2;
  pl.M = a->shape[r - 2];
  pl.K = a->shape[r - 1];
  pl.N = b->shape[r - 1];
  B200_REQUIRE(b->shape[r - 2] == pl.K, B200_ERR_SHAPE, "matmul inner dims differ: lhs K=%lld, rhs K=%lld",
               (long long)pl.K, (long long)b->shape[r - 2]);
  for (int d = 0; d < mm::kMaxBatchDims; ++d) {
    pl.batch[d] = 1;
    pl.a_sb[d] = pl.b_sb[d] = pl.c_sb[d] = 0;
    pl.a_bsz[d] = pl.b_bsz[d] = 1;
  }
  for (int d = 0; d < pl.nb; ++d) {
    const int slot = mm::kMaxBatchDims - pl.nb + d;
    const int64_t da = a->shape[d], db = b->shape[d];
    B200_REQUIRE(da == db || da == 1 || db == 1, B200_ERR_SHAPE, "matmul batch dim %d not broadcastable: %lld vs %lld",
                 d, (long long)da, (long long)db);
    pl.batch[slot] = (int32_t)std::max(da, db);
    pl.a_bsz[slot] = (int32_t)da;
    pl.b_bsz[slot] = (int32_t)db;
    pl.a_sb[slot] = da > 1 ? a->strides[d] : 0;
    pl.b_sb[slot] = db > 1 ? b->strides[d] : 0;
  }
  if (c) {
    B200_REQUIRE(c->rank == r, B200_ERR_SHAPE, "matmul output rank %d != %d", c->rank, r);
    B200_REQUIRE(c->shape[r - 2] == pl.M && c->shape[r - 1] == pl.N, B200_ERR_SHAPE, "matmul output is [%lld, %lld], expected [%lld, %lld]",
                 (long long)c->shape[r - 2], (long long)c->shape[r - 1], (long long)pl.M, (long long)pl.N);
    B200_REQUIRE(c->dtype == B200_F32, B200_ERR_UNSUPPORTED, "matmul output must be f32");
    B200_REQUIRE(c->shape[r - 1] == 1 || c->strides[r - 1] == 1, B200_ERR_UNSUPPORTED, "matmul output rows must be contiguous");
    for (int d = 0; d < pl.nb; ++d) {
      const int slot = mm::kMaxBatchDims - pl.nb + d;
      B200_REQUIRE(c->shape[d] == pl.batch[slot], B200_ERR_SHAPE, "matmul output batch dim %d is %lld, expected %d", d,
                   (long long)c->shape[d], pl.batch[slot]);
      pl.c_sb[slot] = c->strides[d];
    }
  }
  return B200_OK;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bytes of scratch one operand needs under `precision` (0 when it can be consumed in place).
static size_t operand_ws_bytes(const b200_tensor *t, bool is_a, int precision, const MatmulPlan &pl) {
  const int r = t->rank;
  const int64_t mn = is_a ? pl.M : pl.N, k = pl.K;
  int64_t own_batch = 1;
  for (int d = 0; d < pl.nb; ++d) own_batch *= t->shape[d];
  if (precision == B200_MM_F32X3) return align_up((size_t)own_batch * mn * align_up(3 * k, 4) * 4, 256);
  if (precision == B200_MM_BF16) {
    if (t->dtype == B200_BF16) {
      Operand o{t->ptr, 2, is_a ? t->strides[r - 2] : t->strides[r - 1], is_a ? t->strides[r - 1] : t->strides[r - 2], {0, 0, 0}, {1, 1, 1}, false};
      for (int d = 0; d < pl.nb; ++d) {
        o.s_b[mm::kMaxBatchDims - pl.nb + d] = t->strides[d];
        o.bsz[mm::kMaxBatchDims - pl.nb + d] = (int32_t)t->shape[d];
      }
      o.mn_major = o.s_k != 1 && o.s_mn == 1;
      if (tma_ok(o, mn, k)) return 0;
    }
    return align_up((size_t)own_batch * mn * align_up(k, 8) * 2, 256);
  }
  // TF32: in place when TMA can address the f32 view
  Operand o{t->ptr, 4, is_a ? t->strides[r - 2] : t->strides[r - 1], is_a ? t->strides[r - 1] : t->strides[r - 2], {0, 0, 0}, {1, 1, 1}, false};
  for (int d = 0; d < pl.nb; ++d) {
    o.s_b[mm::kMaxBatchDims - pl.nb + d] = t->strides[d];
    o.bsz[mm::kMaxBatchDims - pl.nb + d] = (int32_t)t->shape[d];
  }
  o.mn_major = !(o.s_k == 1 || k == 1) && (o.s_mn == 1 || mn == 1);
  if (t->dtype == B200_F32 && tma_ok(o, mn, k)) return 0;
  return align_up((size_t)own_batch * mn * align_up(k, 4) * 4, 256);
}

template <int ES, bool A_MN, bool B_MN>
static int32_t launch_gemm(const mm::Params &P, const TapeParams &T, size_t epi_bytes, cudaStream_t stream) {
  auto kern = mm::gemm_tcgen05_kernel<ES, A_MN, B_MN>;
  const size_t stage_bytes = 32 * 1024;
  const size_t smem = 1024 + (size_t)P.stages * stage_bytes + 256 + epi_bytes + 64;
  B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int n_tiles = P.tiles_m * P.tiles_n * P.batch[0] * P.batch[1] * P.batch[2];
  const unsigned grid = (unsigned)std::max(1, std::min(n_tiles, sm_count()));
  kern<<<grid, mm::kThreads, smem, stream>>>(P, T);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" int32_t b200_matmul_workspace_bytes(const b200_tensor *a, const b200_tensor *b, int32_t precision,
                                               uint64_t *bytes) {
  B200_REQUIRE(bytes, B200_ERR_INVALID, "bytes is null");
  *bytes = 0;
  B200_REQUIRE(precision >= B200_MM_TF32 && precision <= B200_MM_F32X3, B200_ERR_INVALID, "bad precision %d", precision);
  MatmulPlan pl;
  int32_t st = plan_matmul(a, b, nullptr, pl);
  if (st != B200_OK) return st;
  *bytes = operand_ws_bytes(a, true, precision, pl) + operand_ws_bytes(b, false, precision, pl);
  return B200_OK;
}

extern "C" int32_t b200_launch_matmul(const b200_tensor *a, const b200_tensor *b, const b200_tensor *c,
                                      int32_t precision, const b200_tape *epilogue, const b200_tensor *epi_inputs,
                                      int32_t n_epi_inputs, void *workspace, uint64_t workspace_bytes, b200_stream s) {
  B200_REQUIRE(a && b && c, B200_ERR_INVALID, "null tensor");
  B200_REQUIRE(precision >= B200_MM_TF32 && precision <= B200_MM_F32X3, B200_ERR_INVALID, "bad precision %d", precision);
  B200_REQUIRE(a->dtype == B200_F32 || a->dtype == B200_BF16, B200_ERR_UNSUPPORTED, "matmul lhs dtype %d", a->dtype);
  B200_REQUIRE(b->dtype == a->dtype, B200_ERR_INVALID, "matmul operand dtypes differ");
  B200_REQUIRE(a->dtype == B200_F32 || precision == B200_MM_BF16, B200_ERR_INVALID,
               "bf16 operands require precision B200_MM_BF16");
  MatmulPlan pl;
  int32_t st = plan_matmul(a, b, c, pl);
  if (st != B200_OK) return st;
  const int64_t n_batch = (int64_t)pl.batch[0] * pl.batch[1] * pl.batch[2];
  if (pl.M == 0 || pl.N == 0 || n_batch == 0) return B200_OK;
  cudaStream_t stream = resolve_stream(s);
  const int r = a->rank;

  if (pl.K == 0) {  // empty contraction: C = 0 (then the epilogue would apply; keep it simple)
    B200_REQUIRE(!epilogue, B200_ERR_UNSUPPORTED, "epilogue with K == 0");
    b200_tape_op z = {B200_OP_MOV, (uint8_t)B200_ARG_SCALAR(0), 0, 0, B200_DST_NONE, 0, {0, 0}};
    const uint32_t zero = 0;
    b200_tape tape = {&z, 1, &zero, 1};
    return b200_launch_elemwise(&tape, nullptr, 0, c, 1, c->rank, c->shape, s);
  }

  const size_t need_a = operand_ws_bytes(a, true, precision, pl), need_b = operand_ws_bytes(b, false, precision, pl);
  B200_REQUIRE(need_a + need_b <= workspace_bytes && (need_a + need_b == 0 || workspace), B200_ERR_INVALID,
               "matmul needs %zu bytes of workspace, got %llu", need_a + need_b, (unsigned long long)workspace_bytes);

  // ---- operand views the kernel will consume
  const int es = precision == B200_MM_BF16 ? 2 : 4;
  int64_t K_eff = pl.K;
  auto prepare = [&](const b200_tensor *t, bool is_a, size_t need, char *ws, Operand &o) -> int32_t {
    const int64_t mn = is_a ? pl.M : pl.N;
    o.ptr = t->ptr;
    o.es = es;
    o.s_mn = is_a ? t->strides[r - 2] : t->strides[r - 1];
    o.s_k = is_a ? t->strides[r - 1] : t->strides[r - 2];
    for (int d = 0; d < mm::kMaxBatchDims; ++d) {
      o.s_b[d] = is_a ? pl.a_sb[d] : pl.b_sb[d];
      o.bsz[d] = is_a ? pl.a_bsz[d] : pl.b_bsz[d];
    }
    o.mn_major = !(o.s_k == 1 || pl.K == 1) && (o.s_mn == 1 || mn == 1);
    if (need == 0) return B200_OK;
    // materialise a K-major [own batch..., mn, Kp] copy in the workspace
    int64_t own_batch = 1;
    for (int d = 0; d < mm::kMaxBatchDims; ++d) own_batch *= o.bsz[d];
    if (precision == B200_MM_F32X3) {
      B200_REQUIRE(t->dtype == B200_F32, B200_ERR_UNSUPPORTED, "F32X3 needs f32 operands");
      const int64_t K3p = (int64_t)align_up(3 * pl.K, 4);
      B200_CUDA(cudaMemsetAsync(ws, 0, need, stream));
      mm::SplitBatch sbatch;
      for (int d = 0; d < mm::kMaxBatchDims; ++d) {
        sbatch.bsz[d] = std::max(o.bsz[d], 1);
        sbatch.s_b[d] = o.bsz[d] > 1 ? o.s_b[d] : 0;
      }
      const int64_t total = own_batch * mn * pl.K;
      const unsigned grid = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 16);
      mm::split3_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float *>(t->ptr), o.s_mn, o.s_k, sbatch,
                                                   reinterpret_cast<float *>(ws), (int)mn, (int)pl.K, (int)K3p,
                                                   is_a ? 0 : 1, total);
      B200_LAUNCH_CHECK();
      o.ptr = ws;
      o.s_k = 1;
      o.s_mn = K3p;
      o.mn_major = false;
      int64_t acc = mn * K3p;
      for (int d = mm::kMaxBatchDims - 1; d >= 0; --d) {
        o.s_b[d] = o.bsz[d] > 1 ? acc : 0;
        if (o.bsz[d] > 1) acc *= o.bsz[d];
      }
      return B200_OK;
    }
    // TF32 / BF16: strided copy (with cast) through the elementwise kernel
    const int64_t Kp = (int64_t)align_up(pl.K, es == 2 ? 8 : 4);
    if (Kp != pl.K) B200_CUDA(cudaMemsetAsync(ws, 0, need, stream));
    b200_tensor src, dst;
    src.ptr = t->ptr;
    src.dtype = t->dtype;
    dst.ptr = ws;
    dst.dtype = es == 2 ? B200_BF16 : B200_F32;
    src.rank = dst.rank = mm::kMaxBatchDims + 2;
    int64_t acc = mn * Kp;
    for (int d = mm::kMaxBatchDims - 1; d >= 0; --d) {
      src.shape[d] = dst.shape[d] = o.bsz[d];
      src.strides[d] = o.s_b[d];
      dst.strides[d] = acc;
      acc *= o.bsz[d];
    }
    src.shape[mm::kMaxBatchDims] = dst.shape[mm::kMaxBatchDims] = mn;
    src.shape[mm::kMaxBatchDims + 1] = dst.shape[mm::kMaxBatchDims + 1] = pl.K;
    src.strides[mm::kMaxBatchDims] = o.s_mn;
    src.strides[mm::kMaxBatchDims + 1] = o.s_k;
    dst.strides[mm::kMaxBatchDims] = Kp;
    dst.strides[mm::kMaxBatchDims + 1] = 1;
    int32_t cst = b200_launch_copy(&src, &dst, s);
    if (cst != B200_OK) return cst;
    o.ptr = ws;
    o.s_k = 1;
    o.s_mn = Kp;
    o.mn_major = false;
    for (int d = 0; d < mm::kMaxBatchDims; ++d) o.s_b[d] = o.bsz[d] > 1 ? dst.strides[d] : 0;
    return B200_OK;
  };
  Operand oa, ob;
  st = prepare(a, true, need_a, reinterpret_cast<char *>(workspace), oa);
  if (st != B200_OK) return st;
  st = prepare(b, false, need_b, reinterpret_cast<char *>(workspace) + need_a, ob);
  if (st != B200_OK) return st;
  if (precision == B200_MM_F32X3) K_eff = 3 * pl.K;

  // ---- kernel parameters
  mm::Params P;
  memset(&P, 0, sizeof(P));
  st = make_tmap(&P.tma_a, oa, pl.M, K_eff);
  if (st != B200_OK) return st;
  st = make_tmap(&P.tma_b, ob, pl.N, K_eff);
  if (st != B200_OK) return st;
  P.c = reinterpret_cast<float *>(c->ptr);
  P.ldc = c->strides[r - 2];
  if (pl.M == 1) P.ldc = pl.N;
  P.M = (int32_t)pl.M;
  P.N = (int32_t)pl.N;
  P.K = (int32_t)K_eff;
  for (int d = 0; d < mm::kMaxBatchDims; ++d) {
    P.batch[d] = pl.batch[d];
    P.c_batch_stride[d] = pl.c_sb[d];
    P.a_bflag[d] = oa.bsz[d] > 1 ? 1 : 0;
    P.b_bflag[d] = ob.bsz[d] > 1 ? 1 : 0;
  }
  P.tiles_m = (int32_t)((pl.M + mm::BM - 1) / mm::BM);
  P.tiles_n = (int32_t)((pl.N + mm::BN - 1) / mm::BN);
  const int BK = 128 / es;
  P.k_blocks = (int32_t)((K_eff + BK - 1) / BK);
  P.c_dtype = c->dtype;

  // ---- epilogue tape
  TapeParams T;
  memset(&T, 0, sizeof(T));
  size_t epi_bytes = 0;
  if (epilogue) {
    B200_REQUIRE(n_epi_inputs >= 0 && n_epi_inputs + 1 <= B200_MAX_TAPE_INPUTS, B200_ERR_INVALID, "too many epilogue inputs");
    B200_REQUIRE(pl.N % 4 == 0, B200_ERR_UNSUPPORTED, "a fused matmul epilogue needs N %% 4 == 0 (got N = %lld)", (long long)pl.N);
    CompiledTape ct;
    st = compile_tape(epilogue, n_epi_inputs + 1, 1, ct);
    if (st != B200_OK) return st;
    st = finalize_tape(ct, mm::kEpiU, mm::kEpiBlock, 1, T);
    if (st != B200_OK) return st;
    // describe operands at the collapsed output shape [batch, M, N]
    auto as3 = [&](const b200_tensor &t, OperandDesc &d, const char *what, int idx) -> int32_t {
      B200_REQUIRE(t.rank == r, B200_ERR_SHAPE, "%s %d has rank %d, expected %d", what, idx, t.rank, r);
      int64_t sb = 0, expect = -1;
      for (int dd = pl.nb - 1; dd >= 0; --dd) {
        const int slot = mm::kMaxBatchDims - pl.nb + dd;
        B200_REQUIRE(t.shape[dd] == pl.batch[slot] || t.shape[dd] == 1, B200_ERR_SHAPE, "%s %d: batch dim %d mismatch", what, idx, dd);
        if (pl.batch[slot] <= 1) continue;
        const int64_t sd = t.shape[dd] == 1 ? 0 : t.strides[dd];
        if (expect < 0) { sb = sd; expect = sd * pl.batch[slot]; }
        else { B200_REQUIRE(sd == expect, B200_ERR_UNSUPPORTED, "%s %d: batch dims must be jointly strided", what, idx); expect *= pl.batch[slot]; }
      }
      B200_REQUIRE((t.shape[r - 2] == pl.M || t.shape[r - 2] == 1) && (t.shape[r - 1] == pl.N || t.shape[r - 1] == 1),
                   B200_ERR_SHAPE, "%s %d is not broadcastable to the matmul output", what, idx);
      memset(&d, 0, sizeof(d));
      d.ptr = t.ptr;
      d.dtype = t.dtype;
      d.s3[0] = (int32_t)sb;
      d.s3[1] = (int32_t)(t.shape[r - 2] == 1 ? 0 : t.strides[r - 2]);
      d.s3[2] = (int32_t)(t.shape[r - 1] == 1 ? 0 : t.strides[r - 1]);
      const int esz = dtype_size(t.dtype);
      const bool vec = d.s3[2] == 1 && ((uintptr_t)t.ptr) % (size_t)(esz * 4) == 0 && d.s3[1] % 4 == 0 && d.s3[0] % 4 == 0;
      d.mode = vec ? kModeVec : (d.s3[2] == 0 ? kModeBcast : kModeGather);
      return B200_OK;
    };
    for (int i = 0; i < n_epi_inputs; ++i) {
      st = as3(epi_inputs[i], T.in[i + 1], "epilogue input", i);
      if (st != B200_OK) return st;
    }
    st = as3(*c, T.out[0], "matmul output", 0);
    if (st != B200_OK) return st;
    T.rank = 3;
    P.has_epilogue = 1;
    epi_bytes = slot_file_bytes(T.n_in + T.n_tmp, T.n_scalars, 4, mm::kEpiU, mm::kEpiBlock) + 16;
  }
  const size_t budget = (size_t)max_smem_optin() - 2048 - epi_bytes;
  P.stages = (int32_t)std::max<size_t>(2, std::min<size_t>(6, budget / (32 * 1024)));

  const bool amn = oa.mn_major, bmn = ob.mn_major;
  if (es == 4) {
    if (!amn && !bmn) return launch_gemm<4, false, false>(P, T, epi_bytes, stream);
    if (!amn && bmn) return launch_gemm<4, false, true>(P, T, epi_bytes, stream);
    if (amn && !bmn) return launch_gemm<4, true, false>(P, T, epi_bytes, stream);
    return launch_gemm<4, true, true>(P, T, epi_bytes, stream);
  }
  if (!amn && !bmn) return launch_gemm<2, false, false>(P, T, epi_bytes, stream);
  if (!amn && bmn) return launch_gemm<2, false, true>(P, T, epi_bytes, stream);
  if (amn && !bmn) return launch_gemm<2, true, false>(P, T, epi_bytes, stream);
  return launch_gemm<2, true, true>(P, T, epi_bytes, stream);
}
