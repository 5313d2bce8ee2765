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
#include "tcgen05.cuh"

namespace b200 {
namespace mm {

constexpr int BM = 128, BN = 128;
constexpr int kThreads = 320;            // TMA warp, MMA warp, 8 epilogue warps (4 of them fast-epilogue only)
constexpr int kAccStages = 2;
constexpr int kEpiBlock = 128, kEpiU = 4;  // epilogue tape geometry: 16 columns per dispatch
constexpr int kMaxFastSteps = 6;
constexpr int kRasterGroup = 8;         // m-blocks per rasterisation group (see tile_coords)
constexpr int kOutStageBytes = 8 * 4096;  // 8 epilogue warps x one [32 rows x 128 B] staging buffer

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
  int32_t k_split;   // batch[0] enumerates K splits: k offset = b[0] * k_blocks * BK, partial outputs
  // fast epilogue (epi_fast != 0): the tape is a straight chain acc = OP(acc, operand) evaluated
  // on the 32 accumulator columns a lane holds, staged through swizzled smem and written with TMA
  CUtensorMap tma_c;
  int32_t epi_fast, n_steps;
  struct Step {
    int32_t op;
    int32_t b_kind, b_idx;   // 0 none, 1 input (TapeParams.in[idx]), 3 scalar (bits in b_bits)
    uint32_t b_bits;
    int32_t c_kind, c_idx;
    uint32_t c_bits;
  } steps[kMaxFastSteps];
};

// One operand of a fast-epilogue step for the 32 columns [n0, n0+32) of row m.
__device__ __forceinline__ void epi_fetch(const TapeParams &T, int kind, int idx, uint32_t bits, int batch_lin, int m,
                                          int n0, bool row_ok, int N, uint32_t (&o)[32]) {
  if (kind != 1) {
#pragma unroll
    for (int j = 0; j < 32; ++j) o[j] = bits;
    return;
  }
  const OperandDesc &d = T.in[idx];
  const int64_t base = (int64_t)batch_lin * d.s3[0] + (int64_t)m * d.s3[1];
  if (d.s3[2] == 0) {
    const uint32_t v = row_ok ? load_one(d.ptr, d.dtype, base) : 0u;
#pragma unroll
    for (int j = 0; j < 32; ++j) o[j] = v;
    return;
  }
  const int64_t off = base + (int64_t)n0 * d.s3[2];
  if (d.mode == kModeVec && d.dtype == B200_F32) {
    const uint4 *p4 = reinterpret_cast<const uint4 *>(reinterpret_cast<const uint32_t *>(d.ptr) + off);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      uint4 v = make_uint4(0, 0, 0, 0);
      if (row_ok && n0 + q * 4 < N) v = __ldg(p4 + q);
      o[q * 4] = v.x; o[q * 4 + 1] = v.y; o[q * 4 + 2] = v.z; o[q * 4 + 3] = v.w;
    }
  } else if (d.mode == kModeVec && (d.dtype == B200_BOOL || d.dtype == B200_U8)) {
    const uint32_t *p1 = reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(d.ptr) + off);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint32_t w = (row_ok && n0 + q * 4 < N) ? __ldg(p1 + q) : 0u;
      o[q * 4] = w & 0xFFu; o[q * 4 + 1] = (w >> 8) & 0xFFu; o[q * 4 + 2] = (w >> 16) & 0xFFu; o[q * 4 + 3] = w >> 24;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) o[j] = (row_ok && n0 + j < N) ? load_one(d.ptr, d.dtype, off + (int64_t)j * d.s3[2]) : 0u;
  }
}

// Applies the fast-epilogue chain to a lane's 32 accumulator columns.  Every op rounds exactly as
// the tape interpreter's (tape.cuh run_tape) — same intrinsics, same erf — so fused == unfused.
__device__ __forceinline__ void epi_apply(const Params &P, const TapeParams &T, int batch_lin, int m, int n0, bool row_ok,
                                          uint32_t (&r)[32]) {
  for (int s = 0; s < P.n_steps; ++s) {
    const Params::Step &st = P.steps[s];
    uint32_t b[32];
    if (st.b_kind) epi_fetch(T, st.b_kind, st.b_idx, st.b_bits, batch_lin, m, n0, row_ok, P.N, b);
#define B200_EPI_BIN(expr)                                \
  _Pragma("unroll") for (int j = 0; j < 32; ++j) {        \
    const float x = f_of(r[j]), y = f_of(b[j]);           \
    r[j] = u_of(expr);                                    \
  }
    switch (st.op) {
      case B200_OP_ADD_F: B200_EPI_BIN(__fadd_rn(x, y)) break;
      case B200_OP_SUB_F: B200_EPI_BIN(__fsub_rn(x, y)) break;
      case B200_OP_MUL_F: B200_EPI_BIN(__fmul_rn(x, y)) break;
      case B200_OP_DIV_F: B200_EPI_BIN(__fdiv_rn(x, y)) break;
      case B200_OP_MIN_F: B200_EPI_BIN((x != x || y != y) ? __int_as_float(0x7fc00000) : fminf(x, y)) break;
      case B200_OP_MAX_F: B200_EPI_BIN((x != x || y != y) ? __int_as_float(0x7fc00000) : fmaxf(x, y)) break;
      case kOpDivScalar: {
        const float rinv = __frcp_rn(f_of(b[0]));   // b is the scalar divisor, broadcast
        B200_EPI_BIN((div_scalar_safe(x) ? div_scalar_fast(x, y, rinv) : __fdiv_rn(x, y)))
        break;
      }
      case kOpMulAdd: {
        uint32_t c[32];
        epi_fetch(T, st.c_kind, st.c_idx, st.c_bits, batch_lin, m, n0, row_ok, P.N, c);
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = u_of(__fadd_rn(__fmul_rn(f_of(r[j]), f_of(b[j])), f_of(c[j])));
        break;
      }
      case kOpGelu: {
        const float s2 = 1.41421353816986083984375f, rinv = 0.707106769084930419921875f;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float x = f_of(r[j]);
          const float q = div_scalar_safe(x) ? div_scalar_fast(x, s2, rinv) : __fdiv_rn(x, s2);
          r[j] = u_of(__fmul_rn(__fmul_rn(x, __fadd_rn(erf_f32(q), 1.0f)), 0.5f));
        }
        break;
      }
      case B200_OP_SELECT: {  // acc = C ? B : acc
        uint32_t c[32];
        epi_fetch(T, st.c_kind, st.c_idx, st.c_bits, batch_lin, m, n0, row_ok, P.N, c);
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = c[j] ? b[j] : r[j];
        break;
      }
      default: break;
    }
#undef B200_EPI_BIN
  }
}

// ------------------------------------------------------------------ kernel
// ES = operand element size (4: tf32, 2: bf16).  A_MN / B_MN: operand is MN-major in memory.
// CTAS = 1: one CTA per 128x128 tile.  CTAS = 2: a CTA pair (cluster of 2, cta_group::2) per 256x256
// tile — each CTA stages its own 128 rows of A and 128 of the 256 B rows (same 32 KB per stage for
// twice the math: halves the L2→smem bytes per flop, which is what caps the single-CTA kernel at
// ~50% of the tensor peak), the leader CTA issues M=256,N=256 MMAs that read both CTAs' smem and
// write each CTA's half of the accumulator into its own TMEM; each CTA runs its own epilogue.
template <int ES, bool A_MN, bool B_MN, int CTAS>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ Params P, const __grid_constant__ TapeParams T) {
  constexpr bool PAIR = CTAS == 2;
  constexpr int TN = PAIR ? 256 : BN;     // accumulator columns per tile (per CTA)
  constexpr int BK = 128 / ES;            // K elements per stage (one 128-byte swizzle row)
  constexpr int UMMA_K = 32 / ES;         // K per tcgen05.mma
  constexpr uint32_t kATile = BM * BK * ES, kBTile = BN * BK * ES;  // 16 KB each
  constexpr uint32_t kStageBytes = kATile + kBTile;
  constexpr bool BF16 = ES == 2;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *tiles = smem;
  uint8_t *out_stage = tiles + (size_t)P.stages * kStageBytes;   // 1024-aligned; kOutStageBytes when epi_fast
  uint64_t *full = reinterpret_cast<uint64_t *>(out_stage + (P.epi_fast ? kOutStageBytes : 0));
  uint64_t *empty = full + P.stages;
  uint64_t *acc_full = empty + P.stages;
  uint64_t *acc_empty = acc_full + kAccStages;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + kAccStages);
  uint32_t *epi_smem = reinterpret_cast<uint32_t *>(
      (reinterpret_cast<uintptr_t>(tmem_slot + 4) + 15) & ~uintptr_t(15));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_batch = P.batch[0] * P.batch[1] * P.batch[2];
  const int n_tiles = P.tiles_m * P.tiles_n * n_batch;          // tiles of (128*CTAS) x TN
  const int cta_rank = PAIR ? (int)cluster_ctarank() : 0;
  const int worker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;      // tile-loop start
  const int n_workers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;    // tile-loop stride

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tma_a);
    tma_prefetch_desc(&P.tma_b);
    if (P.epi_fast) tma_prefetch_desc(&P.tma_c);
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], (P.epi_fast ? 8 : 4) * CTAS);  // one arrive per working epilogue warp (of both CTAs in a pair)
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) tmem_alloc_pair(tmem_slot, kAccStages * TN);
    else tmem_alloc(tmem_slot, kAccStages * TN);
  }
  if (P.has_epilogue && !P.epi_fast && warp >= 2 && warp < 6) {
    SlotFile<4, kEpiU, kEpiBlock> slots;
    slots.smem = epi_smem;
    slots.tid = threadIdx.x - 64;
    init_scalars<4, kEpiU, kEpiBlock>(T, slots, T.n_in + T.n_tmp);
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();   // the peer's barriers must be initialised before anything lands on them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Grouped rasterisation: consecutive tile ids (= the tiles one wave of workers runs concurrently) walk
  // kRasterGroup m-blocks down, then one n-block across, so a wave covers a near-square patch of the
  // output and every A / B k-slice it streams is shared by ~8 CTAs through L2 (an m-fastest walk made
  // each wave re-read all of A: 4.4 TB/s of HBM traffic at 16384^3).
  const int per_batch = P.tiles_m * P.tiles_n;
  auto tile_coords = [&](int tile, int &m_blk, int &n_blk, int (&b)[kMaxBatchDims]) {
    int r = tile / per_batch;
    const int t = tile - r * per_batch;
    const int span = kRasterGroup * P.tiles_n;
    const int group = t / span, in_group = t - group * span;
    const int first_m = group * kRasterGroup;
    const int gsz = min(kRasterGroup, P.tiles_m - first_m);
    n_blk = in_group / gsz;
    m_blk = first_m + (in_group - n_blk * gsz);
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
      auto load = [&](void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3, int c4) {
        if constexpr (PAIR) tma_load_5d_pair(dst, map, bar, c0, c1, c2, c3, c4);
        else tma_load_5d(dst, map, bar, c0, c1, c2, c3, c4);
      };
      for (int tile = worker; tile < n_tiles; tile += n_workers) {
        int m_blk, n_blk, b[kMaxBatchDims];
        tile_coords(tile, m_blk, n_blk, b);
        const int ab0 = b[0] * P.a_bflag[0], ab1 = b[1] * P.a_bflag[1], ab2 = b[2] * P.a_bflag[2];
        const int bb0 = b[0] * P.b_bflag[0], bb1 = b[1] * P.b_bflag[1], bb2 = b[2] * P.b_bflag[2];
        const int m0 = (m_blk * CTAS + cta_rank) * BM;           // this CTA's 128 rows of A
        const int n0 = n_blk * TN + cta_rank * BN;               // this CTA's 128 rows of B
        const int kb0 = P.k_split ? b[0] * P.k_blocks : 0;      // split-K: this tile's slice of K
        for (int kbi = 0; kbi < P.k_blocks; ++kbi) {
          const int kb = kb0 + kbi;
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t *sa = tiles + (size_t)stage * kStageBytes;
          uint8_t *sb = sa + kATile;
          // a pair accounts both CTAs' bytes on the leader's barrier
          if (cta_rank == 0) mbar_expect_tx(&full[stage], kStageBytes * CTAS);
          if constexpr (!A_MN) {
            load(sa, &P.tma_a, &full[stage], kb * BK, m0, ab2, ab1, ab0);
          } else {
#pragma unroll
            for (int g = 0; g < BM / BK; ++g)  // BK == elements per 128-byte MN group
              load(sa + g * (BK * 128), &P.tma_a, &full[stage], m0 + g * BK, kb * BK, ab2, ab1, ab0);
          }
          if constexpr (!B_MN) {
            load(sb, &P.tma_b, &full[stage], kb * BK, n0, bb2, bb1, bb0);
          } else {
#pragma unroll
            for (int g = 0; g < BN / BK; ++g)
              load(sb + g * (BK * 128), &P.tma_b, &full[stage], n0 + g * BK, kb * BK, bb2, bb1, bb0);
          }
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1 && cta_rank == 0) {
    // ===================== MMA issuer (the leader CTA of a pair) =====================
    // instruction descriptor: D = F32, A/B = TF32 or BF16, majors, N, M
    uint32_t idesc = 0;
    idesc |= 1u << 4;                              // c_format = F32
    idesc |= (BF16 ? 1u : 2u) << 7;                // a_format
    idesc |= (BF16 ? 1u : 2u) << 10;               // b_format
    idesc |= (A_MN ? 1u : 0u) << 15;               // a_major
    idesc |= (B_MN ? 1u : 0u) << 16;               // b_major
    idesc |= (uint32_t)(TN >> 3) << 17;            // n_dim
    idesc |= (uint32_t)((BM * CTAS) >> 4) << 24;   // m_dim
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = worker; tile < n_tiles; tile += n_workers) {
      if constexpr (PAIR) mbar_wait_cluster(&acc_empty[acc], acc_phase ^ 1);
      else mbar_wait(&acc_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(acc * TN);
      for (int kb = 0; kb < P.k_blocks; ++kb) {
        if constexpr (PAIR) mbar_wait_cluster(&full[stage], phase);
        else mbar_wait(&full[stage], phase);
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
            if constexpr (PAIR) umma_pair<BF16>(tmem_d, da, db, idesc, (kb | k) ? 1u : 0u);
            else umma<BF16>(tmem_d, da, db, idesc, (kb | k) ? 1u : 0u);
          }
        }
        __syncwarp();
        if (lane == 0) {                            // frees the smem stage (in both CTAs) when these MMAs retire
          if constexpr (PAIR) umma_commit_pair(&empty[stage]);
          else umma_commit(&empty[stage]);
        }
        if (++stage == P.stages) { stage = 0; phase ^= 1; }
      }
      if (lane == 0) {                              // accumulator complete → epilogue (of both CTAs)
        if constexpr (PAIR) umma_commit_pair(&acc_full[acc]);
        else umma_commit(&acc_full[acc]);
      }
      __syncwarp();
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 2 && (P.epi_fast || warp < 6)) {
    // ===================== epilogue (warps 2..9; 6..9 only help the fast epilogue) =====================
    // a warp may only touch TMEM lanes 32*(warp%4)..+31: warps w and w+4 share a quarter and take
    // alternate 32-column chunks of it
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row_in_tile = quarter * 32 + lane;
    SlotFile<4, kEpiU, kEpiBlock> slots;
    slots.smem = epi_smem;
    slots.tid = threadIdx.x - 64;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = worker; tile < n_tiles; tile += n_workers) {
      int m_blk, n_blk, b[kMaxBatchDims];
      tile_coords(tile, m_blk, n_blk, b);
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const int m = (m_blk * CTAS + cta_rank) * BM + row_in_tile;
      const int64_t c_batch = b[0] * P.c_batch_stride[0] + b[1] * P.c_batch_stride[1] + b[2] * P.c_batch_stride[2];
      float *crow = P.c + c_batch + (int64_t)m * P.ldc;
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * TN);
      const int batch_lin = (b[0] * P.batch[1] + b[1]) * P.batch[2] + b[2];
      if (P.epi_fast) {
        // 32 columns per pass: TMEM → registers → op chain → 128B-swizzled smem → TMA store (full
        // 128-byte lines, rows/columns past M/N clipped by the tensor map)
        const int m_warp = (m_blk * CTAS + cta_rank) * BM + quarter * 32;
        uint8_t *buf = out_stage + (warp - 2) * 4096;
#pragma unroll 1
        for (int c0 = half * 32; c0 < TN; c0 += 64) {
          const int n0 = n_blk * TN + c0;
          if (n0 >= P.N || m_warp >= P.M) break;
          uint32_t r[32];
          tmem_ld32(taddr + c0, r);
          tmem_ld_wait();
          if (P.n_steps) epi_apply(P, T, batch_lin, m, n0, m < P.M, r);
          if (lane == 0) bulk_wait_read<0>();          // the previous store has finished reading the buffer
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<uint4 *>(buf + lane * 128 + ((q ^ (lane & 7)) << 4)) =
                make_uint4(r[q * 4], r[q * 4 + 1], r[q * 4 + 2], r[q * 4 + 3]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_5d(&P.tma_c, buf, n0, m_warp, b[2], b[1], b[0]);
            bulk_commit();
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (PAIR) mbar_arrive_leader(&acc_empty[acc]);
          else mbar_arrive(&acc_empty[acc]);
        }
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
        continue;
      }
#pragma unroll 1
      for (int c0 = 0; c0 < TN; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(taddr + c0, r);
        tmem_ld_wait();
        const int n0 = n_blk * TN + c0;
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
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_leader(&acc_empty[acc]);
        else mbar_arrive(&acc_empty[acc]);
      }
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
  }

  if (P.epi_fast && warp >= 2 && lane == 0) bulk_wait_all();   // this lane's TMA stores have landed
  __syncwarp();                             // the elected producer / issuer lane rejoins its warp
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();   // neither CTA may exit (or free TMEM) while its peer still uses it
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, kAccStages * TN);
    else tmem_dealloc(tmem_base, kAccStages * TN);
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
static bool tma_c_ok(const b200_tensor *c, int r, const int64_t *c_sb, const int32_t *batch, int64_t M) {
  if (c->dtype != B200_F32 || ((uintptr_t)c->ptr & 15)) return false;
  if (c->strides[r - 1] != 1 && c->shape[r - 1] != 1) return false;
  if (M > 1 && (c->strides[r - 2] % 4 != 0 || c->strides[r - 2] < c->shape[r - 1])) return false;
  for (int d = 0; d < mm::kMaxBatchDims; ++d)
    if (batch[d] > 1 && (c_sb[d] % 4 != 0 || c_sb[d] <= 0)) return false;
  return true;
}

static int32_t make_tmap_c(CUtensorMap *map, void *ptr, int64_t M, int64_t N, int64_t ldc, const int64_t *c_sb,
                           const int32_t *batch) {
  EncodeTiledFn enc = encode_fn();
  B200_REQUIRE(enc, B200_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable from the driver");
  cuuint64_t dims[5], strides[4];
  cuuint32_t box[5] = {32, 32, 1, 1, 1}, estr[5] = {1, 1, 1, 1, 1};
  dims[0] = (cuuint64_t)N;
  dims[1] = (cuuint64_t)M;
  strides[0] = (cuuint64_t)std::max<int64_t>(ldc, 4) * 4;
  for (int d = 0; d < mm::kMaxBatchDims; ++d) {
    const int src = mm::kMaxBatchDims - 1 - d;
    dims[2 + d] = (cuuint64_t)std::max(batch[src], 1);
    strides[1 + d] = (cuuint64_t)(batch[src] > 1 ? c_sb[src] : 4) * 4;
  }
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_REQUIRE(r == CUDA_SUCCESS, B200_ERR_CUDA, "cuTensorMapEncodeTiled (output) failed with %d", (int)r);
  return B200_OK;
}

// Recognises an epilogue tape that is one straight chain acc = OP(acc, operand) ending in the single
// output, with no temporaries — bias add, scaling, mask fill, gelu/relu and their compositions.
static bool match_fast_epilogue(const CompiledTape &ct, const b200_tensor *epi_inputs, mm::Params &P) {
  const size_t n = ct.ops.size();
  if (n < 1 || n - 1 > (size_t)mm::kMaxFastSteps) return false;
  const SymOp &first = ct.ops[0];
  if (first.op != kOpLoad || first.b.kind != 1 || first.b.idx != 0 || first.dst_tmp >= 0) return false;
  if (n == 1) return first.dst_out == 0;
  if (first.dst_out >= 0) return false;
  auto arg_ok = [&](const SymArg &a, bool allow_bool) {
    if (a.kind == 3) return a.idx >= 0 && (size_t)a.idx < ct.scalars.size();
    if (a.kind != 1 || a.idx < 1) return false;
    const int dt = epi_inputs[a.idx - 1].dtype;
    return dt == B200_F32 || (allow_bool && (dt == B200_BOOL || dt == B200_U8));
  };
  P.n_steps = 0;
  for (size_t i = 1; i < n; ++i) {
    const SymOp &o = ct.ops[i];
    if (o.dst_tmp >= 0) return false;
    if ((i + 1 == n) != (o.dst_out == 0) || (i + 1 != n && o.dst_out >= 0)) return false;
    mm::Params::Step st;
    memset(&st, 0, sizeof(st));
    st.op = o.op;
    switch (o.op) {
      case B200_OP_ADD_F: case B200_OP_SUB_F: case B200_OP_MUL_F: case B200_OP_DIV_F:
      case B200_OP_MIN_F: case B200_OP_MAX_F:
        if (!arg_ok(o.b, false)) return false;
        break;
      case kOpDivScalar:
        if (o.b.kind != 3 || !arg_ok(o.b, false)) return false;
        break;
      case kOpMulAdd:
        if (!arg_ok(o.b, false) || !arg_ok(o.c, false)) return false;
        break;
      case kOpGelu:
        if (o.b.kind != 0) return false;   // gelu of the accumulator only
        break;
      case B200_OP_SELECT:
        if (!arg_ok(o.b, false) || !arg_ok(o.c, true)) return false;
        break;
      default: return false;
    }
    auto put = [&](const SymArg &a, int32_t &kind, int32_t &idx, uint32_t &bits) {
      kind = a.kind == 1 ? 1 : (a.kind == 3 ? 3 : 0);
      idx = a.kind == 1 ? a.idx : 0;
      bits = a.kind == 3 ? ct.scalars[a.idx] : 0u;
    };
    put(o.b, st.b_kind, st.b_idx, st.b_bits);
    put(o.c, st.c_kind, st.c_idx, st.c_bits);
    P.steps[P.n_steps++] = st;
  }
  return true;
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

// Tile configuration: CTA pairs (256x256) or single CTAs (128x128), and how many K splits.
// Costs are in units of (one wave) x (K per split): a pair runs a 256x256 tile on two SMs at the
// full tensor rate, a single CTA a 128x128 tile at about half of it, so a wave costs the same.
// Split-K (partials [S, ..., M, N] in the workspace + one deterministic column reduce) fills the
// machine when a long-K product has few output tiles — e.g. a [1024, 8192]·[8192, 1024] weight
// gradient is 16 pair tiles for 74 pairs.
struct GemmConfig {
  bool pair;
  int splits;
};
static GemmConfig choose_config(const MatmulPlan &pl, int es, bool allow_split) {
  static const bool no_pair = std::getenv("B200_MM_NO_PAIR") != nullptr;
  static const bool no_split = std::getenv("B200_MM_NO_SPLITK") != nullptr;
  const int64_t n_batch = (int64_t)pl.batch[0] * pl.batch[1] * pl.batch[2];
  const int64_t sms = sm_count();
  const int64_t t128 = ((pl.M + 127) / 128) * ((pl.N + 127) / 128) * n_batch;
  const int64_t t256 = ((pl.M + 255) / 256) * ((pl.N + 255) / 256) * n_batch;
  const int BK = 128 / es;
  const bool can_split = allow_split && !no_split && pl.nb <= 2 && pl.batch[0] == 1;
  auto splits_for = [&](int64_t units, int64_t capacity) -> int64_t {
    if (!can_split) return 1;
    // each split must keep >= 32 k-blocks: below that the extra partials pass + combine launch cost more
    // than the idle SMs they fill (1024^3 tf32 ran 2.6x slower split in two)
    return std::max<int64_t>(1, std::min<int64_t>({(int64_t)8, capacity / std::max<int64_t>(units, 1), pl.K / (32 * BK)}));
  };
  const bool can_pair = !no_pair && pl.M > 128 && pl.N > 128;
  const int64_t s1 = splits_for(t128, sms), s2 = can_pair ? splits_for(t256, sms / 2) : 1;
  const int64_t cost1 = ((t128 * s1 + sms - 1) / sms) * ((pl.K + s1 - 1) / s1);
  const int64_t cost2 = ((t256 * s2 + sms / 2 - 1) / (sms / 2)) * ((pl.K + s2 - 1) / s2);
  GemmConfig c;
  c.pair = can_pair && (t256 >= sms / 2 || cost2 < cost1);
  c.splits = (int)(c.pair ? s2 : s1);
  return c;
}
static size_t split_ws_bytes(const MatmulPlan &pl, int splits) {
  if (splits <= 1) return 0;
  const int64_t n_batch = (int64_t)pl.batch[0] * pl.batch[1] * pl.batch[2];
  return align_up((size_t)splits * n_batch * pl.M * pl.N * 4, 256);
}

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

template <int ES, bool A_MN, bool B_MN, int CTAS>
static int32_t launch_gemm_n(const mm::Params &P, const TapeParams &T, size_t epi_bytes, cudaStream_t stream) {
  auto kern = mm::gemm_tcgen05_kernel<ES, A_MN, B_MN, CTAS>;
  const size_t stage_bytes = 32 * 1024;
  const size_t smem = 1024 + (size_t)P.stages * stage_bytes + (P.epi_fast ? mm::kOutStageBytes : 0) + 256 + epi_bytes + 64;
  { const int32_t est = ensure_dyn_smem(reinterpret_cast<const void *>(kern), smem); if (est != B200_OK) return est; }
  const int n_tiles = P.tiles_m * P.tiles_n * P.batch[0] * P.batch[1] * P.batch[2];
  if (CTAS == 1) {
    const unsigned grid = (unsigned)std::max(1, std::min(n_tiles, sm_count()));
    kern<<<grid, mm::kThreads, smem, stream>>>(P, T);
  } else {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2u * (unsigned)std::max(1, std::min(n_tiles, sm_count() / 2)));
    cfg.blockDim = dim3(mm::kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    B200_CUDA(cudaLaunchKernelEx(&cfg, kern, P, T));
  }
  B200_LAUNCH_CHECK();
  return B200_OK;
}

template <int ES, bool A_MN, bool B_MN>
static int32_t launch_gemm(mm::Params &P, const TapeParams &T, size_t epi_bytes, cudaStream_t stream, bool pair) {
  if (pair) {
    P.tiles_m = (P.M + 2 * mm::BM - 1) / (2 * mm::BM);
    P.tiles_n = (P.N + 255) / 256;
    return launch_gemm_n<ES, A_MN, B_MN, 2>(P, T, epi_bytes, stream);
  }
  return launch_gemm_n<ES, A_MN, B_MN, 1>(P, T, epi_bytes, stream);
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
  const GemmConfig cfg = choose_config(pl, precision == B200_MM_BF16 ? 2 : 4, true);
  *bytes = operand_ws_bytes(a, true, precision, pl) + operand_ws_bytes(b, false, precision, pl) +
           split_ws_bytes(pl, cfg.splits);
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
  // split-K only for plain f32 outputs without an epilogue, and only when the caller's workspace has
  // room for the partials (b200_matmul_workspace_bytes reports them)
  GemmConfig cfg = choose_config(pl, precision == B200_MM_BF16 ? 2 : 4, !epilogue && c->dtype == B200_F32);
  if (cfg.splits > 1 && need_a + need_b + split_ws_bytes(pl, cfg.splits) > workspace_bytes)
    cfg = choose_config(pl, precision == B200_MM_BF16 ? 2 : 4, false);
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
  float *partials = nullptr;
  if (cfg.splits > 1) {
    // split-K: batch slot 0 enumerates the K slices; each writes its own [batch.., M, N] partial
    partials = reinterpret_cast<float *>(reinterpret_cast<char *>(workspace) + need_a + need_b);
    P.c = partials;
    P.ldc = pl.N;
    P.k_split = 1;
    P.k_blocks = (P.k_blocks + cfg.splits - 1) / cfg.splits;
    P.batch[0] = cfg.splits;
    P.a_bflag[0] = P.b_bflag[0] = 0;
    P.c_batch_stride[2] = pl.M * pl.N;
    P.c_batch_stride[1] = P.c_batch_stride[2] * pl.batch[2];
    P.c_batch_stride[0] = P.c_batch_stride[1] * pl.batch[1];
  }

  // ---- epilogue tape
  TapeParams T;
  memset(&T, 0, sizeof(T));
  size_t epi_bytes = 0;
  // the fast epilogue (TMA stores of full 128-byte lines) needs a 16-byte-aligned f32 output
  // with 16-byte-multiple row and batch strides; B200_MM_LEGACY_EPILOGUE=1 disables it (debug)
  static const bool legacy_epi = std::getenv("B200_MM_LEGACY_EPILOGUE") != nullptr;
  bool fast = !legacy_epi && (partials ? pl.N % 4 == 0 : tma_c_ok(c, r, pl.c_sb, pl.batch, pl.M));
  if (epilogue) {
    B200_REQUIRE(n_epi_inputs >= 0 && n_epi_inputs + 1 <= B200_MAX_TAPE_INPUTS, B200_ERR_INVALID, "too many epilogue inputs");
    B200_REQUIRE(pl.N % 4 == 0, B200_ERR_UNSUPPORTED, "a fused matmul epilogue needs N %% 4 == 0 (got N = %lld)", (long long)pl.N);
    CompiledTape ct;
    st = compile_tape(epilogue, n_epi_inputs + 1, 1, ct);
    if (st != B200_OK) return st;
    fast = fast && match_fast_epilogue(ct, epi_inputs, P);
    if (!fast) P.n_steps = 0;
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
    P.has_epilogue = fast ? 0 : 1;
    if (!fast) epi_bytes = slot_file_bytes(T.n_in + T.n_tmp, T.n_scalars, 4, mm::kEpiU, mm::kEpiBlock) + 16;
  }
  if (fast) {
    st = make_tmap_c(&P.tma_c, P.c, pl.M, pl.N, P.ldc, P.c_batch_stride, P.batch);
    if (st != B200_OK) return st;
    P.epi_fast = 1;
    epi_bytes = 0;
  }
  const size_t budget = (size_t)max_smem_optin() - 2048 - epi_bytes - (fast ? mm::kOutStageBytes : 0);
  P.stages = (int32_t)std::max<size_t>(2, std::min<size_t>(6, budget / (32 * 1024)));

  const bool pair = cfg.pair;
  const bool amn = oa.mn_major, bmn = ob.mn_major;
  if (es == 4) {
    if (!amn && !bmn) st = launch_gemm<4, false, false>(P, T, epi_bytes, stream, pair);
    else if (!amn && bmn) st = launch_gemm<4, false, true>(P, T, epi_bytes, stream, pair);
    else if (amn && !bmn) st = launch_gemm<4, true, false>(P, T, epi_bytes, stream, pair);
    else st = launch_gemm<4, true, true>(P, T, epi_bytes, stream, pair);
  } else {
    if (!amn && !bmn) st = launch_gemm<2, false, false>(P, T, epi_bytes, stream, pair);
    else if (!amn && bmn) st = launch_gemm<2, false, true>(P, T, epi_bytes, stream, pair);
    else if (amn && !bmn) st = launch_gemm<2, true, false>(P, T, epi_bytes, stream, pair);
    else st = launch_gemm<2, true, true>(P, T, epi_bytes, stream, pair);
  }
  if (st != B200_OK || !partials) return st;
  // deterministic combine of the K-split partials: C = sum over the leading axis
  b200_tensor pin, cout;
  memset(&pin, 0, sizeof(pin));
  memset(&cout, 0, sizeof(cout));
  pin.ptr = partials;
  pin.dtype = cout.dtype = B200_F32;
  pin.rank = cout.rank = r + 1;
  cout.ptr = c->ptr;
  int64_t in_shape[B200_MAX_RANK], acc = 1;
  for (int d = r - 1; d >= 0; --d) {
    pin.shape[d + 1] = cout.shape[d + 1] = in_shape[d + 1] = c->shape[d];
    pin.strides[d + 1] = acc;
    cout.strides[d + 1] = c->strides[d];
    acc *= c->shape[d];
  }
  pin.shape[0] = in_shape[0] = cfg.splits;
  pin.strides[0] = acc;
  cout.shape[0] = 1;
  cout.strides[0] = acc;
  return b200_launch_reduce(B200_RED_SUM, 0, r + 1, in_shape, nullptr, &pin, 1, nullptr, nullptr, 0, &cout, 1, s);
}
