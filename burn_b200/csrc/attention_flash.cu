// Flash-style attention for training: forward that keeps only per-row softmax statistics, backward that
// recomputes the weights tile by tile — the [B,H,Sq,Sk] score / weight / dS tensors never exist in HBM
// (SURVEY.md §8(f) row 2).  Semantics are ModuleOps::attention
// (crates/burn-backend/src/backend/ops/modules/base.rs:822-830) as pinned by attention_fallback
// (crates/burn-backend/src/backend/ops/modules/attention.rs:15-90): scores = q·kᵀ·scale, bool / causal
// mask_fill(mask_value), NaN-safe softmax (row max clamped to the most negative finite value, row sum to the
// smallest normal), context = weights·v; the backward is what burn-autodiff's reverse walk over that chain
// yields (matmul / mask_fill / softmax backward), with the mask_fill backward zeroing dS where masked.
//
// Head dim 64, f32 storage, tf32 tcgen05 products, f32 softmax in the base-2 domain (one FFMA + one
// MUFU.EX2 per element).  Three kernels, each a TMA → tcgen05 → TMEM → row-thread pipeline: a producer warp
// issues the TMA loads, converged issuer warps issue the products, the row threads (thread = TMEM lane) do the
// softmax math.  What the measurements on B200 said (scripts/ubench/umma_issue.cu, -DFA_TRACE timelines, ncu
// source pages) and the kernels now do:
//   * a [128 x 64 x 8] tf32 instruction occupies the tensor pipe for 48 cycles with both operands in shared memory
//     (the 6 KB of operand reads, not the math, bound it) and 32 cycles with A in tensor memory; the backward
//     kernels therefore keep P / dS in TMEM — the row threads rewrite the S / dP accumulator tiles in place and the
//     next product reads its A operand from there (no st.shared, no fence.proxy.async, 64 KB less shared memory);
//   * every streamed tile is double-buffered: a single buffer exposes one TMA round trip per key block;
//   * one issuer warp per product, each on its own scheduler, hand-overs through mbarriers (tcgen05.commit tracks
//     the issuing thread's instructions only); the issuers run converged with elect.sync inside the asm;
//   * mbarrier waits sleep in hardware (suspend-time hint) instead of polling, and are bounded (trap, not hang);
//   * warp-uniform fast paths for blocks without any masked / out-of-range element; the softmax scale is applied
//     once per dQ / dK element instead of once per dS element; delta comes from the TMA-staged O tile.
//   * CTAs are dispatched heaviest-first over ALL heads (in-order dispatch = LPT list scheduling under a causal mask);
//     the dQ kernel also exists as a persistent kernel (one CTA per SM, static balanced item deal, everything alive
//     across items) that the host picks for causal masks and short key loops.
//
//   flash_fwd_kernel   CTA = 128 query rows, loop over 64-key blocks.  S_j = Q·K_jᵀ lands in one of two TMEM
//                      buffers (S_{j+1} is issued before the row threads have finished S_j), online softmax:
//                      P_j = 2^(s−m_j) goes to smem as the tf32 A operand, PV_j = P_j·V_j into one of two TMEM
//                      buffers with no accumulation, and the row threads fold it into a register accumulator
//                      O = O·2^(m_{j−1}−m_j) + PV_j while the tensor core already works on block j+1.
//                      Saves stats[row] = (m, 1/l): weights are recomputed as 2^(s−m)·(1/l), the forward's
//                      values (no log-sum-exp cancellation when a row is masked with −1e9).
//                      96 KB smem + 256 TMEM columns: two CTAs per SM.
//   flash_bwd_dq_kernel  CTA = 128 query rows, loop over 64-key blocks: S = Q·K_jᵀ and dP = dO·V_jᵀ (double
//                      buffered in TMEM) → dS = P∘(dP − delta) over dP in place → dQ += dS·K_j accumulated in TMEM,
//                      scaled in the epilogue.  delta = rowsum(dO∘O) is stored into stats[row].z for the second kernel.
//   flash_bwd_dkv_kernel CTA = 128 key rows, loop over 64-query blocks, everything transposed so the row thread
//                      is a key: Sᵀ = K·Q_iᵀ, dPᵀ = V·dO_iᵀ → Pᵀ, dSᵀ in place → dV += Pᵀ·dO_i, dK += dSᵀ·Q_i in
//                      TMEM.  Deterministic (no atomics): replicas stay bit-identical.
// Algorithmic HBM bytes: forward q,k,v,out (+16 B/row); backward q,k,v,out,dO,dq,dk,dv — 4·B·H·S·D·4 and
// 8·B·H·S·D·4 bytes; FLOPs 4·Sq·Sk·D forward, 14·Sq·Sk·D backward per head (7 products, S and dP twice).
#include "tcgen05.cuh"

namespace b200 {
namespace fa {

using namespace mm;

// Every mbarrier wait of these kernels is bounded: a protocol error traps (the launch fails with an error the host sees)
// instead of hanging the device.  The try_wait carries a suspend-time hint, so a waiting thread sleeps in hardware until
// the phase completes instead of polling — polls of the producer / issuer / early row warps compete with the row warps'
// arithmetic for issue slots (a seven-instruction poll loop cost the forward kernel 35 %).  Hint 20 us x 2^19 polls: a
// wedged wait traps after at most ~10 s; a real wait (microseconds) completes inside its first sleep.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .u32 n;\nmov.u32 n, 0;\n"
      "FA_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra FA_WAIT_DONE;\n"
      "add.u32 n, n, 1;\n"
      "setp.lt.u32 p, n, %3;\n"
      "@p bra FA_WAIT_LOOP;\n"
      "trap;\n"
      "FA_WAIT_DONE:\n}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(20000u), "r"(1u << 19)
      : "memory");
}

// The issuer WARP runs converged and every lane executes the same issue code with warp-uniform operands; the single
// issuing lane is chosen by elect.sync INSIDE the asm statement.  Issued from an `if (lane == 0)` region instead, the
// compiler cannot prove the descriptors uniform and wraps every tcgen05.mma in a 16-instruction ELECT / R2UR.BROADCAST /
// BRA.U.ANY waterfall — with [128 x 64 x 8] tf32 instructions (32-48 tensor-pipe cycles each) that single thread's issue
// rate, not the tensor pipe or the softmax arithmetic, bounded all three kernels (ncu: tensor pipe active 20-23 %).
__device__ __forceinline__ void umma_e(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p, q;\nelect.sync _|q, 0xffffffff;\nsetp.ne.b32 p, %4, 0;\n"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_ts_e(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p, q;\nelect.sync _|q, 0xffffffff;\nsetp.ne.b32 p, %4, 0;\n"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void commit_e(uint64_t *bar) {
  asm volatile(
      "{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\n"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(smem_u32(bar))
      : "memory");
}

constexpr int HD = 64;
constexpr int kRows = 128;                      // rows a CTA owns (= TMEM lanes)
constexpr int kCols = 64;                       // columns per loop iteration
constexpr uint32_t kBig = kRows * HD * 4;       // [128 x 64] f32 tile: 32 KB
constexpr uint32_t kSmall = kCols * HD * 4;     // [64 x 64] f32 tile: 16 KB
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kFltMax = 3.402823466e+38f, kMinPos = 1.175494351e-38f;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// K-major [rows x 64] f32 tile = two [rows x 32] 128B-swizzled halves, `half_bytes` apart
__device__ __forceinline__ void load_kmajor(uint8_t *dst, const CUtensorMap *map, uint64_t *bar, int row0, int h, int b,
                                            uint32_t half_bytes) {
  tma_load_5d(dst, map, bar, 0, row0, h, b, 0);
  tma_load_5d(dst + half_bytes, map, bar, 32, row0, h, b, 0);
}
// MN-major [64 k-rows x 64 n] f32 tile: four [32 x 32] boxes (32-byte swizzle atoms)
__device__ __forceinline__ void load_mnmajor(uint8_t *dst, const CUtensorMap *map, uint64_t *bar, int row0, int h, int b) {
#pragma unroll
  for (int kb = 0; kb < 2; ++kb)
#pragma unroll
    for (int g = 0; g < 2; ++g) tma_load_5d(dst + kb * 8192 + g * 4096, map, bar, g * 32, row0 + kb * 32, h, b, 0);
}
// D[128 x 64] (+)= A[128 x 64] · B[64 x 64]ᵀ, A and B K-major tiles (halves a_half / b_half bytes apart)
__device__ __forceinline__ void mma_kk(uint32_t tmem_d, uint32_t a, uint32_t a_half, uint32_t b, uint32_t b_half,
                                       uint32_t idesc, bool accumulate) {
#pragma unroll
  for (int k = 0; k < HD / 8; ++k) {
    const uint64_t da = make_desc(a + (k >> 2) * a_half + (k & 3) * 32, 16, 1024);
    const uint64_t db = make_desc(b + (k >> 2) * b_half + (k & 3) * 32, 16, 1024);
    umma_e(tmem_d, da, db, idesc, (accumulate || k) ? 1u : 0u);
  }
}
// D[128 x 64] (+)= A[128 x 64] (tensor memory: lanes = rows, 64 consecutive columns) · B[64 k x 64 n] (MN-major smem tile)
__device__ __forceinline__ void mma_tmn(uint32_t tmem_d, uint32_t tmem_a, uint32_t b, uint32_t idesc_mn, bool accumulate) {
#pragma unroll
  for (int k = 0; k < kCols / 8; ++k) {
    const uint64_t db = make_desc(b + (k >> 2) * 8192 + (k & 3) * 1024, 4096, 512, 1);
    umma_ts_e(tmem_d, tmem_a + k * 8, db, idesc_mn, (accumulate || k) ? 1u : 0u);
  }
}

// PARTS row threads share a query row, CW = 64 / PARTS score columns each (PARTS x 4 row warps: a TMEM lane quarter is
// reachable from warps w with w % 4 == quarter).  Nothing in the backward couples the columns of a row — m, 1/l and delta
// are per-row constants — so more, narrower row threads only add warps for the schedulers to hide MUFU / TMEM latency with.
__device__ __forceinline__ uint32_t make_idesc() {
  uint32_t idesc = 0;
  idesc |= 1u << 4;                       // D = f32
  idesc |= 2u << 7;                       // A = tf32
  idesc |= 2u << 10;                      // B = tf32
  idesc |= (uint32_t)(kCols >> 3) << 17;  // N = 64
  idesc |= (uint32_t)(kRows >> 4) << 24;  // M = 128
  return idesc;
}

// Key blocks entirely above the causal diagonal are skipped: a filled score contributes 2^(fill - max) = 0
// exactly.  Not when Sq > Sk with a FINITE fill value: the first Sq - Sk query rows then see no key at all, every
// score of theirs equals the fill value and the reference's softmax is uniform over ALL keys (with -inf it is the
// NaN-safe all-zero row, which skipping reproduces).
__device__ __forceinline__ bool causal_skip(int causal, int Sq, int Sk, float mask_value) {
  return causal && !(Sq > Sk && mask_value > -INFINITY);
}

// ------------------------------------------------------------------------------------------ forward
struct FwdParams {
  CUtensorMap tma_q, tma_k, tma_v;
  float *out;
  int64_t o_sb, o_sh, o_ss;
  float4 *stats;                 // [B, H, Sq] (m2, 1/l, delta, -)
  const uint8_t *mask;           // optional bool mask, nonzero = masked
  int64_t m_sb, m_sh, m_ss;
  int32_t B, H, Sq, Sk;
  int32_t causal, q_blocks;
  float scale, mask_value;
};

__global__ void __launch_bounds__(192, 2) flash_fwd_kernel(const __grid_constant__ FwdParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // K_j / V_j double-buffered (buffer j & 1); the weights never touch shared memory: the row threads rewrite the S
  // accumulator tile in tensor memory in place and the PV product reads its A operand from there
  uint8_t *sQ = smem, *sK = sQ + kBig, *sV = sK + 2 * kSmall;
  uint64_t *bars = reinterpret_cast<uint64_t *>(sV + 2 * kSmall);
  uint64_t *bar_q = bars, *bar_k = bars + 1 /*[2]*/, *bar_v = bars + 3 /*[2]*/, *bar_s = bars + 5 /*[2]*/, *bar_p = bars + 7 /*[2]*/,
           *bar_o = bars + 9 /*[2]*/;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int work = blockIdx.x;
  // Launch order = work order, globally: ALL heads' heaviest (latest, under a causal mask) query blocks first, so the
  // hardware's in-order CTA dispatch is longest-processing-time-first list scheduling (per-head interleaving left the
  // SMs that drew a heavy block last ~12 % behind; K/V locality across a head's blocks was measured not to matter)
  const int n_bh = P.B * P.H;
  const int qb = P.q_blocks - 1 - (work / n_bh);
  const int bh = work % n_bh, h = bh % P.H, b = bh / P.H;
  const int q0 = qb * kRows;
  const int nkv_all = (P.Sk + kCols - 1) / kCols;
  int nkv = nkv_all;
  if (causal_skip(P.causal, P.Sq, P.Sk, P.mask_value)) {
    const int last_col = min(P.Sk - 1, q0 + kRows - 1 + (P.Sk - P.Sq));
    nkv = last_col < 0 ? 0 : min(nkv_all, last_col / kCols + 1);
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tma_q);
    tma_prefetch_desc(&P.tma_k);
    tma_prefetch_desc(&P.tma_v);
    for (int i = 0; i < 11; ++i) mbar_init(bars + i, (i == 7 || i == 8) ? 4 : 1);   // bar_p: one arrival per row warp
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;        // S0 / P0 @0, S1 / P1 @64, PV0 @128, PV1 @192

  if (warp == 0) {
    // ===================== TMA producers: lane 0 streams K tiles, lane 1 streams V tiles =====================
    // Two independent streams so that neither load waits behind the other's buffer: K_{j+1} is requested the
    // moment S_j retires (about one block ahead of the row threads), V_{j+1} the moment PV_j retires.
    if (lane == 0 && nkv > 0) {
      mbar_expect_tx(bar_q, kBig);
      load_kmajor(sQ, &P.tma_q, bar_q, q0, h, b, kBig / 2);
      uint32_t ph = 0;
      for (int j = 0; j < nkv; ++j) {
        const int buf = j & 1;
        if (j >= 2) { mbar_wait(bar_s + buf, (ph >> buf) & 1u); ph ^= 1u << buf; }   // S_{j-2} retired → buffer free
        mbar_expect_tx(bar_k + buf, kSmall);
        load_kmajor(sK + buf * kSmall, &P.tma_k, bar_k + buf, j * kCols, h, b, kSmall / 2);
      }
    } else if (lane == 1 && nkv > 0) {
      uint32_t ph = 0;
      for (int j = 0; j < nkv; ++j) {
        const int buf = j & 1;
        if (j >= 2) { mbar_wait(bar_o + buf, (ph >> buf) & 1u); ph ^= 1u << buf; }   // PV_{j-2} retired → buffer free
        mbar_expect_tx(bar_v + buf, kSmall);
        load_mnmajor(sV + buf * kSmall, &P.tma_v, bar_v + buf, j * kCols, h, b);
      }
    }
  } else if (warp == 1) {
    if (nkv > 0) {   // the whole warp, converged: see umma_e
      // ===================== MMA issuer =====================
      // S_{j+1} is issued before the wait for P_j, so the score products stay one block ahead of the row threads.  It
      // overwrites P_{j-1}: the PV product that read it was issued earlier by this same thread (the tensor pipe runs a
      // thread's instructions in order), and the row threads delivered P_{j-1} before that.
      const uint32_t idesc = make_idesc(), idesc_mn = idesc | (1u << 16);
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV);
      uint32_t ph_k = 0, ph_v = 0, ph_p = 0;
      mbar_wait(bar_q, 0);
      mbar_wait(bar_k, 0); ph_k ^= 1u;
      tc_fence_after();
      mma_kk(tmem, aQ, kBig / 2, aK, kSmall / 2, idesc, false);
      commit_e(bar_s);
      for (int j = 0; j < nkv; ++j) {
        const int cur = j & 1, nxt = cur ^ 1;
        if (j + 1 < nkv) {
          mbar_wait(bar_k + nxt, (ph_k >> nxt) & 1u); ph_k ^= 1u << nxt;
          tc_fence_after();
          mma_kk(tmem + nxt * 64u, aQ, kBig / 2, aK + nxt * kSmall, kSmall / 2, idesc, false);
          commit_e(bar_s + nxt);
        }
        mbar_wait(bar_v + cur, (ph_v >> cur) & 1u); ph_v ^= 1u << cur;
        mbar_wait(bar_p + cur, (ph_p >> cur) & 1u); ph_p ^= 1u << cur;   // P_j is in tensor memory; PV_{j-2} has been folded in
        tc_fence_after();
        mma_tmn(tmem + 128u + cur * 64u, tmem + cur * 64u, aV + cur * kSmall, idesc_mn, false);
        commit_e(bar_o + cur);
      }
    }
  } else if (warp >= 2) {
    // ===================== softmax warps: thread = query row = TMEM lane =====================
    const int quarter = warp & 3;
    const int r_in = quarter * 32 + lane, row = q0 + r_in;
    const bool row_ok = row < P.Sq;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int causal_limit = row + (P.Sk - P.Sq);
    const uint8_t *mrow = P.mask ? P.mask + (int64_t)b * P.m_sb + (int64_t)h * P.m_sh + (int64_t)row * P.m_ss : nullptr;
    const float scale2 = P.scale * kLog2e, mask2 = P.mask_value * kLog2e;
    uint32_t ph_s = 0, ph_o = 0;
    float m = -kFltMax, l = 0.0f, alpha_prev = 1.0f;
    float o[HD];
#pragma unroll
    for (int c = 0; c < HD; ++c) o[c] = 0.0f;

    for (int j = 0; j < nkv; ++j) {
      const int cur = j & 1;
      uint32_t mw[16];                                            // this block's mask bytes, fetched before the wait
      if (mrow && row_ok) {
#pragma unroll
        for (int q = 0; q < 16; ++q)
          mw[q] = (j * kCols + q * 4 < P.Sk) ? __ldg(reinterpret_cast<const uint32_t *>(mrow + j * kCols) + q) : 0u;
      }
      mbar_wait(bar_s + cur, (ph_s >> cur) & 1u); ph_s ^= 1u << cur;
      tc_fence_after();
      // Two passes over the S tile in TMEM (reads are ~free next to the exp math) keep 32 instead of 64 score
      // registers live beside the 64-wide output accumulator: pass 1 row max, pass 2 weights.
      const int col0 = j * kCols;
      const uint32_t ts = tmem + (cur ? 64u : 0u) + lane_addr;
      auto load_scores = [&](int kb, float (&s)[32]) {   // scaled, masked scores of columns col0 + kb*32 + [0, 32)
        uint32_t r[32];
        tmem_ld32(ts + kb * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) s[c] = __fmul_rn(__uint_as_float(r[c]), scale2);
        const int c0 = col0 + kb * 32;
        if (mrow && row_ok) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const uint32_t w = mw[kb * 8 + q];
            if (w & 0xFFu) s[q * 4] = mask2;
            if (w & 0xFF00u) s[q * 4 + 1] = mask2;
            if (w & 0xFF0000u) s[q * 4 + 2] = mask2;
            if (w & 0xFF000000u) s[q * 4 + 3] = mask2;
          }
        }
        if (P.causal && c0 + 31 > q0 + (P.Sk - P.Sq)) {   // only chunks crossing the diagonal
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (c0 + c > causal_limit) s[c] = mask2;
        }
        if (c0 + 32 > P.Sk) {
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (c0 + c >= P.Sk) s[c] = -INFINITY;
        }
      };
      // warp-uniform: no mask byte set, no column beyond the causal limit of the warp's first row or past Sk, and a
      // positive scale (so that max(s·scale) = max(s)·scale): the block needs one FMNMX per element for the row max and
      // FFMA + MUFU.EX2 + FADD for the weight — the masked form below spends three times that
      bool any_mask = false;
      if (mrow && row_ok) {
#pragma unroll
        for (int q = 0; q < 16; ++q) any_mask |= mw[q] != 0u;
      }
      const bool fast = scale2 > 0.0f && col0 + kCols <= P.Sk && !(P.causal && col0 + kCols - 1 > q0 + quarter * 32 + (P.Sk - P.Sq)) &&
                        !__any_sync(0xffffffffu, any_mask);
      float bm = -INFINITY;
      if (fast) {
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          uint32_t r[32];
          tmem_ld32(ts + kb * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) bm = fmaxf(bm, __uint_as_float(r[c]));
        }
        bm = __fmul_rn(bm, scale2);
      } else {
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          float s[32];
          load_scores(kb, s);
#pragma unroll
          for (int c = 0; c < 32; ++c) bm = fmaxf(bm, s[c]);
        }
      }
      const float m_new = fmaxf(m, bm);               // m starts at -FLT_MAX: the finfo.min clamp (attention.rs:70-72)
      const float alpha = ex2(m - m_new);
      float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
        float s[32];
        if (fast) {
          uint32_t r[32];
          tmem_ld32(ts + kb * 32, r);
          tmem_ld_wait();
          const float nm = -m_new;
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            s[c] = ex2(__fmaf_rn(__uint_as_float(r[c]), scale2, nm));
            s[c + 1] = ex2(__fmaf_rn(__uint_as_float(r[c + 1]), scale2, nm));
            acc0 += s[c];
            acc1 += s[c + 1];
          }
        } else {
          load_scores(kb, s);
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            s[c] = ex2(s[c] - m_new);
            s[c + 1] = ex2(s[c + 1] - m_new);
            acc0 += s[c];
            acc1 += s[c + 1];
          }
        }
        uint32_t pb[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) pb[c] = __float_as_uint(s[c]);
        tmem_st32(ts + kb * 32, pb);                               // P_j over S_j, in place
      }
      l = l * alpha + (acc0 + acc1);
      m = m_new;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p + cur);
      if (j > 0) {
        // fold PV_{j-1} (other TMEM buffer) into the register accumulator while the tensor core runs block j
        mbar_wait(bar_o + (cur ^ 1), (ph_o >> (cur ^ 1)) & 1u); ph_o ^= 1u << (cur ^ 1);
        tc_fence_after();
        const uint32_t to = tmem + 128u + (cur ? 0u : 64u) + lane_addr;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t r[16];
          tmem_ld16(to + ch * 16, r);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 16; ++c) o[ch * 16 + c] = fmaf(o[ch * 16 + c], alpha_prev, __uint_as_float(r[c]));
        }
        tc_fence_before();
      }
      alpha_prev = alpha;
    }
    const float l_fin = fmaxf(l, kMinPos);                       // min_positive clamp (attention.rs:75-76)
    const float rinv = __frcp_rn(l_fin);
    float *orow = P.out + (int64_t)b * P.o_sb + (int64_t)h * P.o_sh + (int64_t)row * P.o_ss;
    if (nkv > 0) {
      mbar_wait(bar_o + ((nkv - 1) & 1), (ph_o >> ((nkv - 1) & 1)) & 1u);
      tc_fence_after();
      const uint32_t to = tmem + 128u + (((nkv - 1) & 1) ? 64u : 0u) + lane_addr;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t r[16];
        tmem_ld16(to + ch * 16, r);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 16; ++c) o[ch * 16 + c] = __fmul_rn(fmaf(o[ch * 16 + c], alpha_prev, __uint_as_float(r[c])), rinv);
      }
    }
    if (row_ok) {
#pragma unroll
      for (int q = 0; q < 16; ++q)
        reinterpret_cast<float4 *>(orow)[q] = make_float4(o[q * 4], o[q * 4 + 1], o[q * 4 + 2], o[q * 4 + 3]);
      P.stats[((int64_t)b * P.H + h) * P.Sq + row] = make_float4(m, rinv, 0.0f, 0.0f);
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

// ------------------------------------------------------------------------------------------ backward
struct BwdParams {
  // dq kernel: q, d_out boxes of 128 rows; k, v boxes of 64 rows (K-major); k_mn MN-major
  // dkv kernel: k, v boxes of 128 rows; q, d_out boxes of 64 rows (K-major); q_mn, do_mn MN-major
  CUtensorMap tma_q, tma_do, tma_k, tma_v, tma_mn0, tma_mn1;
  const float *o_fwd, *d_out;    // rows read directly for delta
  int64_t of_sb, of_sh, of_ss, do_sb, do_sh, do_ss;
  float *g0, *g1;                // dq | dk, dv
  int64_t g0_sb, g0_sh, g0_ss, g1_sb, g1_sh, g1_ss;
  float4 *stats;
  const uint8_t *mask;
  int64_t m_sb, m_sh, m_ss;
  int32_t B, H, Sq, Sk;
  int32_t causal, blocks;        // row tiles per (b, h)
  float scale, mask_value;
};


#ifdef FA_TRACE
// debug build (-DFA_TRACE): block 0 prints, per role and key block, the SM clock at which each wait completed / each
// product was issued, relative to the start of the role code — the pipeline's actual timeline
#define FA_T(arr, j) do { if (blockIdx.x == 0 && (j) < 16) arr[(j)] = clock64() - fa_t0; } while (0)
#define FA_DUMP(tag, n, a0, a1, a2) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) for (int jj = 0; jj < (n) && jj < 16; ++jj) \
    printf("%s %2d %6lld %6lld %6lld\n", tag, jj, a0[jj], a1[jj], a2[jj]); } while (0)
#else
#define FA_T(arr, j) do { } while (0)
#define FA_DUMP(tag, n, a0, a1, a2) do { } while (0)
#endif
template <int PARTS>
__global__ void __launch_bounds__(128 + 128 * PARTS, 1) flash_bwd_dq_kernel(const __grid_constant__ BwdParams P) {
  constexpr int CW = kCols / PARTS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // Every streamed tile is double-buffered (buffer j & 1) so that the TMA round trip of block j+1 is in flight while
  // block j is multiplied.  dS never touches shared memory: the row threads overwrite the dP accumulator tile in
  // tensor memory in place and the dQ product reads its A operand from there.  sO receives the forward output tile,
  // from which delta = rowsum(dO ∘ O) is computed in shared memory (no per-thread global row walks in the prologue).
  uint8_t *sQ = smem, *sdO = sQ + kBig, *sO = sdO + kBig, *sKk = sO + kBig, *sVk = sKk + 2 * kSmall, *sKmn = sVk + 2 * kSmall;
  float *sDelta = reinterpret_cast<float *>(sKmn + 2 * kSmall);      // [PARTS][128] partial row sums
  uint64_t *bars = reinterpret_cast<uint64_t *>(sDelta + PARTS * kRows);
  uint64_t *bar_q = bars, *bar_of = bars + 1, *bar_kv = bars + 2 /*[2]*/, *bar_mn = bars + 4 /*[2]*/, *bar_s = bars + 6 /*[2]*/,
           *bar_o = bars + 8 /*[2]*/, *bar_p = bars + 10 /*[2]*/, *bar_done = bars + 12;
  // bar_p is per buffer too: a row warp may deliver dS_{j+1} before a slower warp has delivered dS_j (S_{j+1} is issued
  // ahead of the wait for dS_j); on a single barrier that early arrival would complete block j's phase one warp short
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int work = blockIdx.x;
  const int n_bh = P.B * P.H;                                        // heaviest blocks of all heads first (see flash_fwd_kernel)
  const int qb = P.blocks - 1 - (work / n_bh);
  const int bh = work % n_bh, h = bh % P.H, b = bh / P.H;
  const int q0 = qb * kRows;
  const int nkv_all = (P.Sk + kCols - 1) / kCols;
  int nkv = nkv_all;
  if (causal_skip(P.causal, P.Sq, P.Sk, P.mask_value)) {
    const int last_col = min(P.Sk - 1, q0 + kRows - 1 + (P.Sk - P.Sq));
    nkv = last_col < 0 ? 0 : min(nkv_all, last_col / kCols + 1);
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tma_q);
    tma_prefetch_desc(&P.tma_do);
    tma_prefetch_desc(&P.tma_k);
    tma_prefetch_desc(&P.tma_v);
    tma_prefetch_desc(&P.tma_mn0);
    tma_prefetch_desc(&P.tma_mn1);
    for (int i = 0; i < 13; ++i)   // bar_s: both the S and the dP issuer commit; bar_p: one arrival per row warp
      mbar_init(bars + i, (i == 10 || i == 11) ? 4 * PARTS : ((i == 6 || i == 7) ? 2 : 1));
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;        // S0 @0, S1 @64, dP0 / dS0 @128, dP1 / dS1 @192, dQ @256
#ifdef FA_TRACE
  const long long fa_t0 = clock64();
  long long tr0[16], tr1[16], tr2[16];
  for (int i = 0; i < 16; ++i) tr0[i] = tr1[i] = tr2[i] = 0;
#endif

  if (warp == 0) {
    // TMA producers: lane 0 streams the K-major K_j / V_j tiles (buffer free when S_{j-2} / dP_{j-2} retired), lane 1
    // the MN-major K_j tile (buffer free when the dQ product of block j-2 retired)
    if (lane == 0 && nkv > 0) {
      mbar_expect_tx(bar_of, 2 * kBig);                                // dO and O first: delta is the row threads' prologue
      load_kmajor(sdO, &P.tma_do, bar_of, q0, h, b, kBig / 2);
      load_kmajor(sO, &P.tma_mn1, bar_of, q0, h, b, kBig / 2);
      mbar_expect_tx(bar_q, kBig);
      load_kmajor(sQ, &P.tma_q, bar_q, q0, h, b, kBig / 2);
      uint32_t ph[2] = {0, 0};
      for (int j = 0; j < nkv; ++j) {
        const int buf = j & 1;
        if (j >= 2) { mbar_wait(bar_s + buf, ph[buf]); ph[buf] ^= 1; }
        mbar_expect_tx(bar_kv + buf, 2 * kSmall);
        load_kmajor(sKk + buf * kSmall, &P.tma_k, bar_kv + buf, j * kCols, h, b, kSmall / 2);
        load_kmajor(sVk + buf * kSmall, &P.tma_v, bar_kv + buf, j * kCols, h, b, kSmall / 2);
      }
    } else if (lane == 1 && nkv > 0) {
      uint32_t ph[2] = {0, 0};
      for (int j = 0; j < nkv; ++j) {
        const int buf = j & 1;
        if (j >= 2) { mbar_wait(bar_o + buf, ph[buf]); ph[buf] ^= 1; }
        mbar_expect_tx(bar_mn + buf, kSmall);
        load_mnmajor(sKmn + buf * kSmall, &P.tma_mn0, bar_mn + buf, j * kCols, h, b);
      }
    }
  } else if (warp <= 3) {
    // Three issuer warps, one product each, each on its own scheduler: with [128 x 64 x 8] tf32 instructions the issue
    // path of a single thread (descriptor arithmetic + five R2UR per tcgen05.mma, ~120 cycles each against 32-48 cycles
    // of tensor-pipe work) bounded the kernel at a quarter of the tensor rate.  tcgen05.commit tracks the issuing
    // thread's own instructions, so every hand-over between the issuers goes through an mbarrier:
    //   S_j   overwrites S_{j-2}   -> the row threads have read it               (bar_p of block j-2)
    //   dP_j  overwrites dS_{j-2}  -> the dQ product of block j-2 has retired    (bar_o of block j-2)
    //   dQ_j  reads dS_j           -> the row threads have delivered it          (bar_p of block j)
    if (nkv > 0) {   // whole warps, converged: see umma_e
      const uint32_t idesc = make_idesc(), idesc_mn = idesc | (1u << 16);
      uint32_t ph_a = 0, ph_b = 0;           // bit b: phase of the per-buffer barrier pair this warp waits on
      if (warp == 1) {
        const uint32_t aQ = smem_u32(sQ), aKk = smem_u32(sKk);
        mbar_wait(bar_q, 0);
        for (int j = 0; j < nkv; ++j) {
          const int buf = j & 1;
          mbar_wait(bar_kv + buf, (ph_a >> buf) & 1u); ph_a ^= 1u << buf;
          FA_T(tr0, j);
          if (j >= 2) { mbar_wait(bar_p + buf, (ph_b >> buf) & 1u); ph_b ^= 1u << buf; }
          FA_T(tr1, j);
          tc_fence_after();
          mma_kk(tmem + buf * 64u, aQ, kBig / 2, aKk + buf * kSmall, kSmall / 2, idesc, false);           // S = Q·Kᵀ
          commit_e(bar_s + buf);
          FA_T(tr2, j);
        }
        FA_DUMP("S  kv/p/issued", nkv, tr0, tr1, tr2);
      } else if (warp == 2) {
        const uint32_t adO = smem_u32(sdO), aVk = smem_u32(sVk);
        mbar_wait(bar_of, 0);
        for (int j = 0; j < nkv; ++j) {
          const int buf = j & 1;
          mbar_wait(bar_kv + buf, (ph_a >> buf) & 1u); ph_a ^= 1u << buf;
          FA_T(tr0, j);
          if (j >= 2) { mbar_wait(bar_o + buf, (ph_b >> buf) & 1u); ph_b ^= 1u << buf; }
          FA_T(tr1, j);
          tc_fence_after();
          mma_kk(tmem + 128u + buf * 64u, adO, kBig / 2, aVk + buf * kSmall, kSmall / 2, idesc, false);   // dP = dO·Vᵀ
          commit_e(bar_s + buf);
          FA_T(tr2, j);
        }
        FA_DUMP("dP kv/o/issued", nkv, tr0, tr1, tr2);
      } else {
        const uint32_t aKmn = smem_u32(sKmn);
        for (int j = 0; j < nkv; ++j) {
          const int cur = j & 1;
          mbar_wait(bar_mn + cur, (ph_a >> cur) & 1u); ph_a ^= 1u << cur;
          FA_T(tr0, j);
          mbar_wait(bar_p + cur, (ph_b >> cur) & 1u); ph_b ^= 1u << cur;   // dS_j is in tensor memory (over dP_j)
          FA_T(tr1, j);
          tc_fence_after();
          mma_tmn(tmem + 256u, tmem + 128u + cur * 64u, aKmn + cur * kSmall, idesc_mn, j > 0);   // dQ += dS_j·K_j
          commit_e(bar_o + cur);
          FA_T(tr2, j);
        }
        FA_DUMP("dQ mn/p/issued", nkv, tr0, tr1, tr2);
        commit_e(bar_done);
      }
    }
  } else {
    const int quarter = warp & 3, part = (warp - 4) >> 2;
    const int r_in = quarter * 32 + lane, row = q0 + r_in;
    const bool row_ok = row < P.Sq;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int causal_limit = row + (P.Sk - P.Sq);
    const uint8_t *mrow = P.mask ? P.mask + (int64_t)b * P.m_sb + (int64_t)h * P.m_sh + (int64_t)row * P.m_ss : nullptr;
    const float scale2 = P.scale * kLog2e, mask2 = P.mask_value * kLog2e;
    float4 *stat = P.stats + ((int64_t)b * P.H + h) * P.Sq + row;
    float nm2 = 0.0f, rinv = 0.0f, delta = 0.0f;
    float4 st4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row_ok) st4 = *stat;                                         // in flight while the tiles land
    if (nkv > 0) {
      // delta = sum_d dO[row, d] * O[row, d]  (= rowsum(dP ∘ P)) from the two tiles in shared memory: this thread sums
      // its half of the row (conflict-free: the 128-byte swizzle spreads 8 consecutive rows over all banks), the halves
      // meet through sDelta.  Rows past Sq are zero-filled by TMA.
      mbar_wait(bar_of, 0);
      const int c_first = part * CW;                                 // this thread's columns of the [128 x 64] tiles
      const uint8_t *pd = sdO + (c_first >> 5) * (kBig / 2) + r_in * 128, *po = sO + (c_first >> 5) * (kBig / 2) + r_in * 128;
      float psum = 0.0f;
#pragma unroll
      for (int qq = 0; qq < CW / 4; ++qq) {
        const int q = ((c_first & 31) >> 2) + qq;
        const int off = (q ^ (r_in & 7)) << 4;
        const float4 g = *reinterpret_cast<const float4 *>(pd + off), ov = *reinterpret_cast<const float4 *>(po + off);
        psum = __fadd_rn(psum, __fadd_rn(__fadd_rn(__fmul_rn(ov.x, g.x), __fmul_rn(ov.y, g.y)),
                                         __fadd_rn(__fmul_rn(ov.z, g.z), __fmul_rn(ov.w, g.w))));
      }
      sDelta[part * kRows + r_in] = psum;
      asm volatile("bar.sync 1, %0;" ::"n"(128 * PARTS) : "memory");   // the row warps: all partial sums are visible
#pragma unroll
      for (int pp = 0; pp < PARTS; ++pp) delta = __fadd_rn(delta, sDelta[pp * kRows + r_in]);
    }
    if (row_ok) {
      nm2 = -st4.x;
      rinv = st4.y;
      if (part == 0) stat->z = delta;      // the dK/dV kernel (launched after this one) reads it per query column
    }
    uint32_t ph_s = 0;
    for (int j = 0; j < nkv; ++j) {
      const int cur = j & 1;
      const int col0 = j * kCols + part * CW;
      uint32_t masked = 0;                                         // bit c: column col0 + c is masked
      if (mrow && row_ok) {                                        // fetched before the wait: latency overlaps it
#pragma unroll
        for (int q = 0; q < CW / 4; ++q) {
          if (col0 + q * 4 < P.Sk) {
            const uint32_t mw = __ldg(reinterpret_cast<const uint32_t *>(mrow + col0) + q);
            if (mw & 0xFFu) masked |= 1u << (q * 4);
            if (mw & 0xFF00u) masked |= 2u << (q * 4);
            if (mw & 0xFF0000u) masked |= 4u << (q * 4);
            if (mw & 0xFF000000u) masked |= 8u << (q * 4);
          }
        }
      }
      if (P.causal && col0 + CW - 1 > causal_limit) {
        const int first = causal_limit + 1 - col0;                 // first masked column of this chunk
        masked |= first <= 0 ? 0xFFFFFFFFu : (first >= 32 ? 0u : (0xFFFFFFFFu << first));
      }
      // warp-uniform: nothing masked in this [32 rows x CW columns] chunk and no column past Sk — 5 instructions per
      // element instead of the masked form's dozen
      const bool fast = (col0 + CW <= P.Sk) && !__any_sync(0xffffffffu, masked != 0u);
      FA_T(tr0, j);
      mbar_wait(bar_s + cur, (ph_s >> cur) & 1u); ph_s ^= 1u << cur;
      FA_T(tr1, j);
      tc_fence_after();
      const uint32_t t_dp = tmem + 128u + cur * 64u + lane_addr + part * CW;
      uint32_t rs[CW], rp[CW];
      tmem_ldn(tmem + cur * 64u + lane_addr + part * CW, rs);
      tmem_ldn(t_dp, rp);
      tmem_ld_wait();
      // dS WITHOUT the softmax scale: dQ = scale · Σ_j dS_j·K_j is scaled once per output element in the epilogue
      if (fast) {
#pragma unroll
        for (int c = 0; c < CW; ++c) {
          const float p = __fmul_rn(ex2(__fmaf_rn(__uint_as_float(rs[c]), scale2, nm2)), rinv);
          rp[c] = __float_as_uint(__fmul_rn(p, __fsub_rn(__uint_as_float(rp[c]), delta)));
        }
      } else {
#pragma unroll
        for (int c = 0; c < CW; ++c) {
          const bool mk = (masked >> c) & 1u;
          const float t = mk ? __fadd_rn(mask2, nm2) : __fmaf_rn(__uint_as_float(rs[c]), scale2, nm2);
          float p = __fmul_rn(ex2(t), rinv);
          if (col0 + c >= P.Sk) p = 0.0f;
          const float d = __fmul_rn(p, __fsub_rn(__uint_as_float(rp[c]), delta));
          rp[c] = __float_as_uint(mk ? 0.0f : d);                   // mask_fill backward: no gradient through a filled score
        }
      }
      tmem_stn(t_dp, rp);                                          // dS_j over dP_j, in place
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p + cur);
      FA_T(tr2, j);
    }
    if (warp == 4) FA_DUMP("row top/s/arrived", nkv, tr0, tr1, tr2);
    float *grow_out = P.g0 + (int64_t)b * P.g0_sb + (int64_t)h * P.g0_sh + (int64_t)row * P.g0_ss + part * CW;
    if (nkv > 0) {
      mbar_wait(bar_done, 0);                                       // every product has retired: dQ is complete
      tc_fence_after();
      uint32_t r[CW];
      tmem_ldn(tmem + 256u + lane_addr + part * CW, r);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int q = 0; q < CW / 4; ++q)
          reinterpret_cast<float4 *>(grow_out)[q] =
              make_float4(__fmul_rn(__uint_as_float(r[q * 4]), P.scale), __fmul_rn(__uint_as_float(r[q * 4 + 1]), P.scale),
                          __fmul_rn(__uint_as_float(r[q * 4 + 2]), P.scale), __fmul_rn(__uint_as_float(r[q * 4 + 3]), P.scale));
      }
    } else if (row_ok) {
#pragma unroll
      for (int q = 0; q < CW / 4; ++q) reinterpret_cast<uint4 *>(grow_out)[q] = make_uint4(0, 0, 0, 0);
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// (The row-thread block bodies of the persistent kernels repeat those of the per-item kernels on purpose: factored into
//  shared device functions — tried — the per-item dK/dV kernel compiled 10 % slower, 140 -> 155 us, from a different
//  register allocation around the 64 live score registers.)
// ---------------------------------------------------------------------------------- persistent dQ kernel
// One CTA per SM walks a STATIC, balanced list of (batch, head, query block) items: the items sorted by work (under a
// causal mask the late query blocks see the most keys), dealt to the CTAs in snake order — for [8,16,1024,64] causal the
// heaviest CTA gets 64 key blocks against a mean of 62.3.  Barriers, the tensor-memory allocation and the tile rings
// live for the whole kernel and every role runs the same item sequence, so the K/V prefetch, the first S / dP products
// and the Q / dO / O loads of item n+1 overlap the last blocks and the dQ read-out of item n.  With one CTA per
// (b, h, block) the per-CTA fixed cost (190 KB of prologue loads under load, allocation, the final barrier, the launch
// order's imbalance) weighs most under a causal mask.  Measured on [8,16,1024,64]: causal 104 -> 94 us, unmasked 132 -> 137 us;
// on the encoder's [64,8,256,64]: backward 162 -> 154 us.  B200_FA_PERSIST=0 selects the one-CTA-per-item kernel.
struct DqItem {
  int q0, h, b, nkv;
  bool valid;
};
__device__ __forceinline__ DqItem dq_item(const BwdParams &P, int round) {
  DqItem it;
  const int G = (int)gridDim.x, c = (int)blockIdx.x;
  const int n_bh = P.B * P.H, total = n_bh * P.blocks;
  // (a head-major order — the eight blocks of a head on neighbouring CTAs, for L2 locality of K / V — was measured: no
  //  faster without a mask, 12 % slower with the causal one, whose imbalance it brings back)
  const int pos = round * G + ((round & 1) ? G - 1 - c : c);
  it.valid = pos < total;
  const int level = it.valid ? pos / n_bh : 0;               // 0 = heaviest
  const int bh = it.valid ? pos - level * n_bh : 0;
  const int qb = P.blocks - 1 - level;
  it.q0 = qb * kRows;
  it.h = bh % P.H;
  it.b = bh / P.H;
  const int nkv_all = (P.Sk + kCols - 1) / kCols;
  it.nkv = nkv_all;
  if (causal_skip(P.causal, P.Sq, P.Sk, P.mask_value)) {
    const int last_col = min(P.Sk - 1, it.q0 + kRows - 1 + (P.Sk - P.Sq));
    it.nkv = last_col < 0 ? 0 : min(nkv_all, last_col / kCols + 1);
  }
  return it;
}

template <int PARTS>
__global__ void __launch_bounds__(128 + 128 * PARTS, 1) flash_bwd_dq_persist_kernel(const __grid_constant__ BwdParams P) {
  constexpr int CW = kCols / PARTS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *sQ = smem, *sdO = sQ + kBig, *sO = sdO + kBig, *sKk = sO + kBig, *sVk = sKk + 2 * kSmall, *sKmn = sVk + 2 * kSmall;
  float *sDelta = reinterpret_cast<float *>(sKmn + 2 * kSmall);      // [PARTS][128] partial row sums
  uint64_t *bars = reinterpret_cast<uint64_t *>(sDelta + PARTS * kRows);
  // per item: bar_q, bar_do, bar_ot (tiles landed), bar_ofree (rows are done with the dO / O tiles: delta computed),
  // bar_done (all dQ products of the item retired), bar_dqfree (rows have read the dQ accumulator);
  // per key block, by TMEM buffer / tile stage g & 1 (g counts key blocks across items): bar_kv, bar_mn, bar_s, bar_o, bar_p
  uint64_t *bar_q = bars, *bar_do = bars + 1, *bar_ot = bars + 2, *bar_ofree = bars + 3, *bar_done = bars + 4, *bar_dqfree = bars + 5,
           *bar_kv = bars + 6 /*[2]*/, *bar_mn = bars + 8 /*[2]*/, *bar_s = bars + 10 /*[2]*/, *bar_o = bars + 12 /*[2]*/,
           *bar_p = bars + 14 /*[2]*/;
  constexpr int kBars = 16;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + kBars);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_items = P.B * P.H * P.blocks;
  const int rounds = (total_items + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tma_q);
    tma_prefetch_desc(&P.tma_do);
    tma_prefetch_desc(&P.tma_k);
    tma_prefetch_desc(&P.tma_v);
    tma_prefetch_desc(&P.tma_mn0);
    tma_prefetch_desc(&P.tma_mn1);
    for (int i = 0; i < kBars; ++i) {
      uint32_t count = 1;
      if (i == 3 || i == 5 || i >= 14) count = 4 * PARTS;    // bar_ofree, bar_dqfree, bar_p: one arrival per row warp
      else if (i == 10 || i == 11) count = 2;                // bar_s: the S and the dP issuer commit
      mbar_init(bars + i, count);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;        // S0 @0, S1 @64, dP0 / dS0 @128, dP1 / dS1 @192, dQ @256

  if (warp == 0) {
    if (lane == 0) {
      // K-major stream (K_g, V_g into stage g & 1: free when S_{g-2} and dP_{g-2} retired) + the per-item tiles
      uint32_t ph_s = 0, ph_of = 0;
      int s_waited = 0;                    // bar_s completions consumed so far (key blocks 0 .. s_waited-1, in order)
      auto wait_s_upto = [&](int x) {      // the S and dP products of key block x (and all earlier ones) have retired
        while (s_waited <= x) {
          const int bf = s_waited & 1;
          mbar_wait(bar_s + bf, (ph_s >> bf) & 1u);
          ph_s ^= 1u << bf;
          ++s_waited;
        }
      };
      int g = 0, n = 0;
      for (int r = 0; r < rounds; ++r) {
        const DqItem it = dq_item(P, r);
        if (!it.valid) break;
        if (it.nkv == 0) continue;
        for (int j = 0; j < it.nkv; ++j) {
          const int gb = g + j, buf = gb & 1;
          if (gb >= 2) wait_s_upto(gb - 2);
          mbar_expect_tx(bar_kv + buf, 2 * kSmall);
          load_kmajor(sKk + buf * kSmall, &P.tma_k, bar_kv + buf, j * kCols, it.h, it.b, kSmall / 2);
          load_kmajor(sVk + buf * kSmall, &P.tma_v, bar_kv + buf, j * kCols, it.h, it.b, kSmall / 2);
          if (j == 0) {
            // the item's own tiles, after its first K/V block (whose stage frees earlier): sQ / sdO were read by the S /
            // dP products of the previous item (all retired once block g-1 has), sdO / sO by its row threads (delta)
            if (n > 0) {
              wait_s_upto(g - 1);
              mbar_wait(bar_ofree, ph_of);
              ph_of ^= 1u;
            }
            mbar_expect_tx(bar_q, kBig);
            load_kmajor(sQ, &P.tma_q, bar_q, it.q0, it.h, it.b, kBig / 2);
            mbar_expect_tx(bar_do, kBig);
            load_kmajor(sdO, &P.tma_do, bar_do, it.q0, it.h, it.b, kBig / 2);
            mbar_expect_tx(bar_ot, kBig);
            load_kmajor(sO, &P.tma_mn1, bar_ot, it.q0, it.h, it.b, kBig / 2);
          }
        }
        g += it.nkv;
        ++n;
      }
    } else if (lane == 1) {
      // MN-major K stream (stage g & 1: free when the dQ product of block g-2 retired)
      uint32_t ph = 0;
      int g = 0;
      for (int r = 0; r < rounds; ++r) {
        const DqItem it = dq_item(P, r);
        if (!it.valid) break;
        for (int j = 0; j < it.nkv; ++j) {
          const int gb = g + j, buf = gb & 1;
          if (gb >= 2) { mbar_wait(bar_o + buf, (ph >> buf) & 1u); ph ^= 1u << buf; }
          mbar_expect_tx(bar_mn + buf, kSmall);
          load_mnmajor(sKmn + buf * kSmall, &P.tma_mn0, bar_mn + buf, j * kCols, it.h, it.b);
        }
        g += it.nkv;
      }
    }
  } else if (warp <= 3) {
    // issuer warps (whole warps, converged: see umma_e), one product each
    const uint32_t idesc = make_idesc(), idesc_mn = idesc | (1u << 16);
    uint32_t ph_a = 0, ph_b = 0, ph_item = 0;
    int g = 0, n = 0;
    for (int r = 0; r < rounds; ++r) {
      const DqItem it = dq_item(P, r);
      if (!it.valid) break;
      if (it.nkv == 0) continue;
      if (warp == 1) {
        const uint32_t aQ = smem_u32(sQ), aKk = smem_u32(sKk);
        mbar_wait(bar_q, ph_item);
        for (int j = 0; j < it.nkv; ++j) {
          const int gb = g + j, buf = gb & 1;
          mbar_wait(bar_kv + buf, (ph_a >> buf) & 1u); ph_a ^= 1u << buf;
          if (gb >= 2) { mbar_wait(bar_p + buf, (ph_b >> buf) & 1u); ph_b ^= 1u << buf; }   // rows have read S_{g-2}
          tc_fence_after();
          mma_kk(tmem + buf * 64u, aQ, kBig / 2, aKk + buf * kSmall, kSmall / 2, idesc, false);           // S = Q·Kᵀ
          commit_e(bar_s + buf);
        }
      } else if (warp == 2) {
        const uint32_t adO = smem_u32(sdO), aVk = smem_u32(sVk);
        mbar_wait(bar_do, ph_item);
        for (int j = 0; j < it.nkv; ++j) {
          const int gb = g + j, buf = gb & 1;
          mbar_wait(bar_kv + buf, (ph_a >> buf) & 1u); ph_a ^= 1u << buf;
          if (gb >= 2) { mbar_wait(bar_o + buf, (ph_b >> buf) & 1u); ph_b ^= 1u << buf; }   // dQ_{g-2} has read dS_{g-2}
          tc_fence_after();
          mma_kk(tmem + 128u + buf * 64u, adO, kBig / 2, aVk + buf * kSmall, kSmall / 2, idesc, false);   // dP = dO·Vᵀ
          commit_e(bar_s + buf);
        }
      } else {
        const uint32_t aKmn = smem_u32(sKmn);
        if (n > 0) mbar_wait(bar_dqfree, ph_item ^ 1u);               // the rows have read the previous item's dQ
        for (int j = 0; j < it.nkv; ++j) {
          const int gb = g + j, cur = gb & 1;
          mbar_wait(bar_mn + cur, (ph_a >> cur) & 1u); ph_a ^= 1u << cur;
          mbar_wait(bar_p + cur, (ph_b >> cur) & 1u); ph_b ^= 1u << cur;   // dS_g is in tensor memory (over dP_g)
          tc_fence_after();
          mma_tmn(tmem + 256u, tmem + 128u + cur * 64u, aKmn + cur * kSmall, idesc_mn, j > 0);   // dQ += dS·K
          commit_e(bar_o + cur);
        }
        commit_e(bar_done);
      }
      g += it.nkv;
      ++n;
      ph_item ^= 1u;
    }
  } else {
    const int quarter = warp & 3, part = (warp - 4) >> 2;
    const int r_in = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const float scale2 = P.scale * kLog2e, mask2 = P.mask_value * kLog2e;
    uint32_t ph_s = 0, ph_item = 0;
    int g = 0;
    for (int r = 0; r < rounds; ++r) {
      const DqItem it = dq_item(P, r);
      if (!it.valid) break;
      const int row = it.q0 + r_in;
      const bool row_ok = row < P.Sq;
      float *grow_out = P.g0 + (int64_t)it.b * P.g0_sb + (int64_t)it.h * P.g0_sh + (int64_t)row * P.g0_ss + part * CW;
      if (it.nkv == 0) {
        if (row_ok) {
#pragma unroll
          for (int q = 0; q < CW / 4; ++q) reinterpret_cast<uint4 *>(grow_out)[q] = make_uint4(0, 0, 0, 0);
        }
        continue;
      }
      const int causal_limit = row + (P.Sk - P.Sq);
      const uint8_t *mrow = P.mask ? P.mask + (int64_t)it.b * P.m_sb + (int64_t)it.h * P.m_sh + (int64_t)row * P.m_ss : nullptr;
      float4 *stat = P.stats + ((int64_t)it.b * P.H + it.h) * P.Sq + row;
      float nm2 = 0.0f, rinv = 0.0f, delta = 0.0f;
      float4 st4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row_ok) st4 = *stat;                                       // in flight while the tiles land
      {
        // delta = sum_d dO[row, d] * O[row, d] from the two tiles in shared memory (see flash_bwd_dq_kernel)
        mbar_wait(bar_do, ph_item);
        mbar_wait(bar_ot, ph_item);
        const int c_first = part * CW;
        const uint8_t *pd = sdO + (c_first >> 5) * (kBig / 2) + r_in * 128, *po = sO + (c_first >> 5) * (kBig / 2) + r_in * 128;
        float psum = 0.0f;
#pragma unroll
        for (int qq = 0; qq < CW / 4; ++qq) {
          const int q = ((c_first & 31) >> 2) + qq;
          const int off = (q ^ (r_in & 7)) << 4;
          const float4 gg = *reinterpret_cast<const float4 *>(pd + off), ov = *reinterpret_cast<const float4 *>(po + off);
          psum = __fadd_rn(psum, __fadd_rn(__fadd_rn(__fmul_rn(ov.x, gg.x), __fmul_rn(ov.y, gg.y)),
                                           __fadd_rn(__fmul_rn(ov.z, gg.z), __fmul_rn(ov.w, gg.w))));
        }
        sDelta[part * kRows + r_in] = psum;
        asm volatile("bar.sync 1, %0;" ::"n"(128 * PARTS) : "memory");   // the row warps: all partial sums are visible
#pragma unroll
        for (int pp = 0; pp < PARTS; ++pp) delta = __fadd_rn(delta, sDelta[pp * kRows + r_in]);
        if (lane == 0) mbar_arrive(bar_ofree);                       // this warp no longer reads the dO / O tiles
      }
      if (row_ok) {
        nm2 = -st4.x;
        rinv = st4.y;
        if (part == 0) stat->z = delta;    // the dK/dV kernel (launched after this one) reads it per query column
      }
      for (int j = 0; j < it.nkv; ++j) {
        const int cur = (g + j) & 1;
        const int col0 = j * kCols + part * CW;
        uint32_t masked = 0;                                         // bit c: column col0 + c is masked
        if (mrow && row_ok) {
#pragma unroll
          for (int q = 0; q < CW / 4; ++q) {
            if (col0 + q * 4 < P.Sk) {
              const uint32_t mw = __ldg(reinterpret_cast<const uint32_t *>(mrow + col0) + q);
              if (mw & 0xFFu) masked |= 1u << (q * 4);
              if (mw & 0xFF00u) masked |= 2u << (q * 4);
              if (mw & 0xFF0000u) masked |= 4u << (q * 4);
              if (mw & 0xFF000000u) masked |= 8u << (q * 4);
            }
          }
        }
        if (P.causal && col0 + CW - 1 > causal_limit) {
          const int first = causal_limit + 1 - col0;                 // first masked column of this chunk
          masked |= first <= 0 ? 0xFFFFFFFFu : (first >= 32 ? 0u : (0xFFFFFFFFu << first));
        }
        const bool fast = (col0 + CW <= P.Sk) && !__any_sync(0xffffffffu, masked != 0u);
        mbar_wait(bar_s + cur, (ph_s >> cur) & 1u); ph_s ^= 1u << cur;
        tc_fence_after();
        const uint32_t t_dp = tmem + 128u + cur * 64u + lane_addr + part * CW;
        uint32_t rs[CW], rp[CW];
        tmem_ldn(tmem + cur * 64u + lane_addr + part * CW, rs);
        tmem_ldn(t_dp, rp);
        tmem_ld_wait();
        if (fast) {
#pragma unroll
          for (int c = 0; c < CW; ++c) {
            const float p = __fmul_rn(ex2(__fmaf_rn(__uint_as_float(rs[c]), scale2, nm2)), rinv);
            rp[c] = __float_as_uint(__fmul_rn(p, __fsub_rn(__uint_as_float(rp[c]), delta)));
          }
        } else {
#pragma unroll
          for (int c = 0; c < CW; ++c) {
            const bool mk = (masked >> c) & 1u;
            const float t = mk ? __fadd_rn(mask2, nm2) : __fmaf_rn(__uint_as_float(rs[c]), scale2, nm2);
            float p = __fmul_rn(ex2(t), rinv);
            if (col0 + c >= P.Sk) p = 0.0f;
            const float d = __fmul_rn(p, __fsub_rn(__uint_as_float(rp[c]), delta));
            rp[c] = __float_as_uint(mk ? 0.0f : d);                 // mask_fill backward: no gradient through a filled score
          }
        }
        tmem_stn(t_dp, rp);                                         // dS over dP, in place
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p + cur);
      }
      {
        mbar_wait(bar_done, ph_item);                               // every dQ product of the item has retired
        tc_fence_after();
        uint32_t rq[CW];
        tmem_ldn(tmem + 256u + lane_addr + part * CW, rq);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_dqfree);                     // the accumulator may be overwritten by the next item
        if (row_ok) {
#pragma unroll
          for (int q = 0; q < CW / 4; ++q)
            reinterpret_cast<float4 *>(grow_out)[q] =
                make_float4(__fmul_rn(__uint_as_float(rq[q * 4]), P.scale), __fmul_rn(__uint_as_float(rq[q * 4 + 1]), P.scale),
                            __fmul_rn(__uint_as_float(rq[q * 4 + 2]), P.scale), __fmul_rn(__uint_as_float(rq[q * 4 + 3]), P.scale));
        }
      }
      g += it.nkv;
      ph_item ^= 1u;
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

constexpr int kDkvThreads = 13 * 32;   // TMA warp, four issuer warps (one product each), 8 row warps

__global__ void __launch_bounds__(kDkvThreads, 1) flash_bwd_dkv_kernel(const __grid_constant__ BwdParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // K, V resident; the four streamed tiles of a query block (Q, dO in both operand layouts) double-buffered (buffer
  // it & 1).  Pᵀ and dSᵀ never touch shared memory: the row threads rewrite the Sᵀ / dPᵀ accumulator tiles in tensor
  // memory in place and the dV / dK products read their A operand from there.
  uint8_t *sK = smem, *sV = sK + kBig, *sQk = sV + kBig, *sdOk = sQk + 2 * kSmall, *sQmn = sdOk + 2 * kSmall, *sdOmn = sQmn + 2 * kSmall;
  // [4][64] (m2, 1/l, delta, -) of the block's queries.  Four buffers: the producer refills buffer it % 4 once
  // Sᵀ_{it-2} has retired, and that product is only issued after the dV / dK products of block it-4 — issued after
  // the row threads delivered block it-4, the last reader of the buffer — have retired.
  float4 *sStats = reinterpret_cast<float4 *>(sdOmn + 2 * kSmall);
  uint64_t *bars = reinterpret_cast<uint64_t *>(sStats + 4 * kCols);
  uint64_t *bar_k = bars, *bar_v = bars + 1, *bar_qk = bars + 2 /*[2]*/, *bar_mn = bars + 4 /*[2]*/, *bar_s = bars + 6 /*[2]*/,
           *bar_o = bars + 8 /*[2]*/, *bar_p = bars + 10 /*[2]*/, *bar_done = bars + 12, *bar_st = bars + 13 /*[4]*/;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 17);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int work = blockIdx.x;
  const int n_bh = P.B * P.H;             // early key tiles see the most queries under a causal mask: all heads' first
  const int kb_ = work / n_bh;
  const int bh = work % n_bh, h = bh % P.H, b = bh / P.H;
  const int kv0 = kb_ * kRows;
  const int nq = (P.Sq + kCols - 1) / kCols;
  int i_start = 0;
  if (causal_skip(P.causal, P.Sq, P.Sk, P.mask_value)) {
    const int first_q = kv0 - (P.Sk - P.Sq);                        // first query row that sees key kv0
    i_start = first_q <= 0 ? 0 : min(nq, first_q / kCols);
  }
  const int n_it = nq - i_start;
  const int64_t stat_base = ((int64_t)b * P.H + h) * P.Sq;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tma_q);
    tma_prefetch_desc(&P.tma_do);
    tma_prefetch_desc(&P.tma_k);
    tma_prefetch_desc(&P.tma_v);
    tma_prefetch_desc(&P.tma_mn0);
    tma_prefetch_desc(&P.tma_mn1);
    for (int i = 0; i < 17; ++i)   // bar_s / bar_o / bar_done: two issuers commit; bar_p: one arrival per row warp
      mbar_init(bars + i, (i == 10 || i == 11) ? 8 : ((i >= 6 && i <= 9) || i == 12 ? 2 : 1));
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;        // ST0 / PT0 @0, ST1 / PT1 @64, dPT0 / dST0 @128, dPT1 / dST1 @192, dV @256, dK @320

  if (warp == 0) {
    // TMA producers: lane 0 streams the K-major Q_i / dO_i tiles + the block's row statistics (buffer free when Sᵀ / dPᵀ
    // of block i-2 retired), lane 1 the MN-major Q_i / dO_i tiles (buffer free when the dV / dK products of block i-2 retired)
    if (lane == 0 && n_it > 0) {
      mbar_expect_tx(bar_k, kBig);
      load_kmajor(sK, &P.tma_k, bar_k, kv0, h, b, kBig / 2);
      mbar_expect_tx(bar_v, kBig);
      load_kmajor(sV, &P.tma_v, bar_v, kv0, h, b, kBig / 2);
      uint32_t ph = 0;
      for (int it = 0; it < n_it; ++it) {
        const int buf = it & 1;
        if (it >= 2) { mbar_wait(bar_s + buf, (ph >> buf) & 1u); ph ^= 1u << buf; }
        const int r0 = (i_start + it) * kCols;
        mbar_expect_tx(bar_qk + buf, 2 * kSmall);
        load_kmajor(sQk + buf * kSmall, &P.tma_q, bar_qk + buf, r0, h, b, kSmall / 2);
        load_kmajor(sdOk + buf * kSmall, &P.tma_do, bar_qk + buf, r0, h, b, kSmall / 2);
        // the block's per-query statistics: one bulk copy, clipped at the end of the sequence
        const uint32_t bytes = (uint32_t)min(kCols, P.Sq - r0) * 16u;
        uint64_t *bst = bar_st + (it & 3);
        mbar_expect_tx(bst, bytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(sStats + (it & 3) * kCols)),
                     "l"(P.stats + stat_base + r0), "r"(bytes), "r"(smem_u32(bst))
                     : "memory");
      }
    } else if (lane == 1 && n_it > 0) {
      uint32_t ph = 0;
      for (int it = 0; it < n_it; ++it) {
        const int buf = it & 1;
        if (it >= 2) { mbar_wait(bar_o + buf, (ph >> buf) & 1u); ph ^= 1u << buf; }
        const int r0 = (i_start + it) * kCols;
        mbar_expect_tx(bar_mn + buf, 2 * kSmall);
        load_mnmajor(sQmn + buf * kSmall, &P.tma_mn0, bar_mn + buf, r0, h, b);
        load_mnmajor(sdOmn + buf * kSmall, &P.tma_mn1, bar_mn + buf, r0, h, b);
      }
    }
  } else if (warp <= 4) {
    // Four issuer warps, one product each (see flash_bwd_dq_kernel).  Hand-overs:
    //   Sᵀ_i  overwrites Pᵀ_{i-2},  dPᵀ_i overwrites dSᵀ_{i-2}  -> the dV / dK products of block i-2 retired  (bar_o)
    //   dV_i reads Pᵀ_i, dK_i reads dSᵀ_i                        -> the row threads delivered them            (bar_p)
    if (n_it > 0) {   // whole warps, converged: see umma_e
      const uint32_t idesc = make_idesc(), idesc_mn = idesc | (1u << 16);
      uint32_t ph_a = 0, ph_b = 0;
      if (warp <= 2) {
        const bool is_s = warp == 1;
        const uint32_t aA = smem_u32(is_s ? sK : sV), aB = smem_u32(is_s ? sQk : sdOk);
        const uint32_t td = tmem + (is_s ? 0u : 128u);
        mbar_wait(is_s ? bar_k : bar_v, 0);
        for (int it = 0; it < n_it; ++it) {
          const int buf = it & 1;
          mbar_wait(bar_qk + buf, (ph_a >> buf) & 1u); ph_a ^= 1u << buf;
          if (it >= 2) { mbar_wait(bar_o + buf, (ph_b >> buf) & 1u); ph_b ^= 1u << buf; }
          tc_fence_after();
          mma_kk(td + buf * 64u, aA, kBig / 2, aB + buf * kSmall, kSmall / 2, idesc, false);   // Sᵀ = K·Qᵀ | dPᵀ = V·dOᵀ
          commit_e(bar_s + buf);
        }
      } else {
        const bool is_v = warp == 3;
        const uint32_t aB = smem_u32(is_v ? sdOmn : sQmn);
        const uint32_t td = tmem + (is_v ? 256u : 320u), ta = tmem + (is_v ? 0u : 128u);
        for (int it = 0; it < n_it; ++it) {
          const int cur = it & 1;
          mbar_wait(bar_mn + cur, (ph_a >> cur) & 1u); ph_a ^= 1u << cur;
          mbar_wait(bar_p + cur, (ph_b >> cur) & 1u); ph_b ^= 1u << cur;    // Pᵀ_i and dSᵀ_i are in tensor memory
          tc_fence_after();
          mma_tmn(td, ta + cur * 64u, aB + cur * kSmall, idesc_mn, it > 0);   // dV += Pᵀ·dO_i | dK += dSᵀ·Q_i
          commit_e(bar_o + cur);
        }
        commit_e(bar_done);
      }
    }
  } else {
    const int quarter = warp & 3, half = (warp - 5) >> 2;
    const int r_in = quarter * 32 + lane, kv = kv0 + r_in;
    const bool kv_ok = kv < P.Sk;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int shift = P.Sk - P.Sq;
    const uint8_t *mcol = P.mask ? P.mask + (int64_t)b * P.m_sb + (int64_t)h * P.m_sh + kv : nullptr;
    const float scale2 = P.scale * kLog2e, mask2 = P.mask_value * kLog2e;
    uint32_t ph_s = 0;
    for (int it = 0; it < n_it; ++it) {
      const int cur = it & 1, sb = it & 3;
      const int qc0 = (i_start + it) * kCols + half * 32;
      uint32_t mbits = 0;                                           // bit c: explicit mask at (query qc0 + c, key kv)
      if (mcol && kv_ok) {                                          // fetched before the waits
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (qc0 + c < P.Sq && __ldg(mcol + (int64_t)(qc0 + c) * P.m_ss)) mbits |= 1u << c;
      }
      // warp-uniform: every (key, query) pair of this [32 keys x 32 queries] chunk is in range and unmasked — 6
      // instructions per element instead of the masked form's two dozen.
      // dSᵀ is delivered WITHOUT the softmax scale: dK = scale · Σ_i dSᵀ_i·Q_i is scaled once per element in the epilogue.
      const bool fast = (qc0 + 32 <= P.Sq) && (kv0 + quarter * 32 + 32 <= P.Sk) &&
                        !(P.causal && kv0 + quarter * 32 + 31 > qc0 + shift) && !__any_sync(0xffffffffu, mbits != 0u);
      mbar_wait(bar_st + sb, (uint32_t)((it >> 2) & 1));
      mbar_wait(bar_s + cur, (ph_s >> cur) & 1u); ph_s ^= 1u << cur;
      tc_fence_after();
      const uint32_t t_s = tmem + cur * 64u + lane_addr + half * 32, t_dp = tmem + 128u + cur * 64u + lane_addr + half * 32;
      uint32_t rs[32], rp[32];
      tmem_ld32(t_s, rs);
      tmem_ld32(t_dp, rp);
      tmem_ld_wait();
      const float4 *st = sStats + sb * kCols + half * 32;
      if (fast) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float4 sq = st[c];                                  // smem broadcast
          const float p = __fmul_rn(ex2(__fmaf_rn(__uint_as_float(rs[c]), scale2, -sq.x)), sq.y);
          rs[c] = __float_as_uint(p);
          rp[c] = __float_as_uint(__fmul_rn(p, __fsub_rn(__uint_as_float(rp[c]), sq.z)));
        }
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const int q = qc0 + c;
          const bool q_ok = q < P.Sq;
          const float4 sq = st[c];                                  // smem broadcast; garbage past Sq is never used
          const bool mk = (P.causal && kv > q + shift) || ((mbits >> c) & 1u);
          const float t = mk ? __fsub_rn(mask2, sq.x) : __fmaf_rn(__uint_as_float(rs[c]), scale2, -sq.x);
          const float p = (q_ok && kv_ok) ? __fmul_rn(ex2(t), sq.y) : 0.0f;
          const float d = __fmul_rn(p, __fsub_rn(__uint_as_float(rp[c]), sq.z));
          rs[c] = __float_as_uint(p);
          rp[c] = __float_as_uint((mk || !(q_ok && kv_ok)) ? 0.0f : d);
        }
      }
      tmem_st32(t_s, rs);                                           // Pᵀ_i over Sᵀ_i, dSᵀ_i over dPᵀ_i, in place
      tmem_st32(t_dp, rp);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p + cur);
    }
    float *dk_row = P.g0 + (int64_t)b * P.g0_sb + (int64_t)h * P.g0_sh + (int64_t)kv * P.g0_ss + half * 32;
    float *dv_row = P.g1 + (int64_t)b * P.g1_sb + (int64_t)h * P.g1_sh + (int64_t)kv * P.g1_ss + half * 32;
    if (n_it > 0) {
      mbar_wait(bar_done, 0);                                       // every product has retired: dV and dK are complete
      tc_fence_after();
      uint32_t rv[32], rk[32];
      tmem_ld32(tmem + 256u + lane_addr + half * 32, rv);
      tmem_ld32(tmem + 320u + lane_addr + half * 32, rk);
      tmem_ld_wait();
      if (kv_ok) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          reinterpret_cast<uint4 *>(dv_row)[q] = make_uint4(rv[q * 4], rv[q * 4 + 1], rv[q * 4 + 2], rv[q * 4 + 3]);
          reinterpret_cast<float4 *>(dk_row)[q] =
              make_float4(__fmul_rn(__uint_as_float(rk[q * 4]), P.scale), __fmul_rn(__uint_as_float(rk[q * 4 + 1]), P.scale),
                          __fmul_rn(__uint_as_float(rk[q * 4 + 2]), P.scale), __fmul_rn(__uint_as_float(rk[q * 4 + 3]), P.scale));
        }
      }
    } else if (kv_ok) {
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        reinterpret_cast<uint4 *>(dv_row)[q] = make_uint4(0, 0, 0, 0);
        reinterpret_cast<uint4 *>(dk_row)[q] = make_uint4(0, 0, 0, 0);
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ---------------------------------------------------------------------------------- persistent dK/dV kernel
// The dK/dV kernel with the item loop of flash_bwd_dq_persist_kernel: items (batch, head, key block) sorted by work
// (under a causal mask the EARLY key blocks see the most queries), dealt to one CTA per SM in snake order; barriers,
// tensor memory and the tile / statistics rings alive across items.
struct DkvItem {
  int kv0, h, b, i_start, n_it;
  bool valid;
};
__device__ __forceinline__ DkvItem dkv_item(const BwdParams &P, int round) {
  DkvItem it;
  const int G = (int)gridDim.x, c = (int)blockIdx.x;
  const int n_bh = P.B * P.H, total = n_bh * P.blocks;
  const int pos = round * G + ((round & 1) ? G - 1 - c : c);
  it.valid = pos < total;
  const int level = it.valid ? pos / n_bh : 0;               // 0 = first key block = heaviest
  const int bh = it.valid ? pos - level * n_bh : 0;
  it.kv0 = level * kRows;
  it.h = bh % P.H;
  it.b = bh / P.H;
  const int nq = (P.Sq + kCols - 1) / kCols;
  it.i_start = 0;
  if (causal_skip(P.causal, P.Sq, P.Sk, P.mask_value)) {
    const int first_q = it.kv0 - (P.Sk - P.Sq);                     // first query row that sees key kv0
    it.i_start = first_q <= 0 ? 0 : min(nq, first_q / kCols);
  }
  it.n_it = nq - it.i_start;
  return it;
}

__global__ void __launch_bounds__(kDkvThreads, 1) flash_bwd_dkv_persist_kernel(const __grid_constant__ BwdParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *sK = smem, *sV = sK + kBig, *sQk = sV + kBig, *sdOk = sQk + 2 * kSmall, *sQmn = sdOk + 2 * kSmall, *sdOmn = sQmn + 2 * kSmall;
  float4 *sStats = reinterpret_cast<float4 *>(sdOmn + 2 * kSmall);   // [4][64] (m2, 1/l, delta, -), ring over query blocks
  uint64_t *bars = reinterpret_cast<uint64_t *>(sStats + 4 * kCols);
  // per item: bar_k, bar_v (tiles landed), bar_done (the item's dV / dK products retired), bar_accfree (rows have read
  // the accumulators); per query block g (counted across items): bar_qk, bar_mn, bar_s, bar_o, bar_p by g & 1, bar_st by g & 3
  uint64_t *bar_k = bars, *bar_v = bars + 1, *bar_done = bars + 2, *bar_accfree = bars + 3, *bar_qk = bars + 4 /*[2]*/,
           *bar_mn = bars + 6 /*[2]*/, *bar_s = bars + 8 /*[2]*/, *bar_o = bars + 10 /*[2]*/, *bar_p = bars + 12 /*[2]*/,
           *bar_st = bars + 14 /*[4]*/;
  constexpr int kBars = 18;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + kBars);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_items = P.B * P.H * P.blocks;
  const int rounds = (total_items + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tma_q);
    tma_prefetch_desc(&P.tma_do);
    tma_prefetch_desc(&P.tma_k);
    tma_prefetch_desc(&P.tma_v);
    tma_prefetch_desc(&P.tma_mn0);
    tma_prefetch_desc(&P.tma_mn1);
    for (int i = 0; i < kBars; ++i) {
      uint32_t count = 1;
      if (i == 3 || i == 12 || i == 13) count = 8;                   // bar_accfree, bar_p: one arrival per row warp
      else if (i == 2 || (i >= 8 && i <= 11)) count = 2;            // bar_done, bar_s, bar_o: two issuers commit
      mbar_init(bars + i, count);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;        // ST0 / PT0 @0, ST1 / PT1 @64, dPT0 / dST0 @128, dPT1 / dST1 @192, dV @256, dK @320

  if (warp == 0) {
    if (lane == 0) {
      uint32_t ph_s = 0;
      int s_waited = 0;
      auto wait_s_upto = [&](int x) {      // Sᵀ and dPᵀ of query block x (and all earlier ones) have retired
        while (s_waited <= x) {
          const int bf = s_waited & 1;
          mbar_wait(bar_s + bf, (ph_s >> bf) & 1u);
          ph_s ^= 1u << bf;
          ++s_waited;
        }
      };
      int g = 0, n = 0;
      for (int r = 0; r < rounds; ++r) {
        const DkvItem it = dkv_item(P, r);
        if (!it.valid) break;
        if (it.n_it == 0) continue;
        const int64_t stat_base = ((int64_t)it.b * P.H + it.h) * P.Sq;
        for (int i = 0; i < it.n_it; ++i) {
          const int gb = g + i, buf = gb & 1;
          if (gb >= 2) wait_s_upto(gb - 2);
          const int r0 = (it.i_start + i) * kCols;
          mbar_expect_tx(bar_qk + buf, 2 * kSmall);
          load_kmajor(sQk + buf * kSmall, &P.tma_q, bar_qk + buf, r0, it.h, it.b, kSmall / 2);
          load_kmajor(sdOk + buf * kSmall, &P.tma_do, bar_qk + buf, r0, it.h, it.b, kSmall / 2);
          const uint32_t bytes = (uint32_t)min(kCols, P.Sq - r0) * 16u;
          uint64_t *bst = bar_st + (gb & 3);
          mbar_expect_tx(bst, bytes);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           smem_u32(sStats + (gb & 3) * kCols)),
                       "l"(P.stats + stat_base + r0), "r"(bytes), "r"(smem_u32(bst))
                       : "memory");
          if (i == 0) {
            // the item's K and V tiles: read by the Sᵀ / dPᵀ products of the previous item, all retired with block g-1
            if (n > 0) wait_s_upto(g - 1);
            mbar_expect_tx(bar_k, kBig);
            load_kmajor(sK, &P.tma_k, bar_k, it.kv0, it.h, it.b, kBig / 2);
            mbar_expect_tx(bar_v, kBig);
            load_kmajor(sV, &P.tma_v, bar_v, it.kv0, it.h, it.b, kBig / 2);
          }
        }
        g += it.n_it;
        ++n;
      }
    } else if (lane == 1) {
      uint32_t ph = 0;
      int g = 0;
      for (int r = 0; r < rounds; ++r) {
        const DkvItem it = dkv_item(P, r);
        if (!it.valid) break;
        for (int i = 0; i < it.n_it; ++i) {
          const int gb = g + i, buf = gb & 1;
          if (gb >= 2) { mbar_wait(bar_o + buf, (ph >> buf) & 1u); ph ^= 1u << buf; }
          const int r0 = (it.i_start + i) * kCols;
          mbar_expect_tx(bar_mn + buf, 2 * kSmall);
          load_mnmajor(sQmn + buf * kSmall, &P.tma_mn0, bar_mn + buf, r0, it.h, it.b);
          load_mnmajor(sdOmn + buf * kSmall, &P.tma_mn1, bar_mn + buf, r0, it.h, it.b);
        }
        g += it.n_it;
      }
    }
  } else if (warp <= 4) {
    const uint32_t idesc = make_idesc(), idesc_mn = idesc | (1u << 16);
    uint32_t ph_a = 0, ph_b = 0, ph_item = 0;
    int g = 0, n = 0;
    const bool first_pair = warp <= 2, is_s = warp == 1, is_v = warp == 3;
    const uint32_t aA = smem_u32(is_s ? sK : sV), aBk = smem_u32(is_s ? sQk : sdOk), aBmn = smem_u32(is_v ? sdOmn : sQmn);
    for (int r = 0; r < rounds; ++r) {
      const DkvItem it = dkv_item(P, r);
      if (!it.valid) break;
      if (it.n_it == 0) continue;
      if (first_pair) {
        mbar_wait(is_s ? bar_k : bar_v, ph_item);
        const uint32_t td = tmem + (is_s ? 0u : 128u);
        for (int i = 0; i < it.n_it; ++i) {
          const int gb = g + i, buf = gb & 1;
          mbar_wait(bar_qk + buf, (ph_a >> buf) & 1u); ph_a ^= 1u << buf;
          if (gb >= 2) { mbar_wait(bar_o + buf, (ph_b >> buf) & 1u); ph_b ^= 1u << buf; }   // dV / dK of block g-2 retired
          tc_fence_after();
          mma_kk(td + buf * 64u, aA, kBig / 2, aBk + buf * kSmall, kSmall / 2, idesc, false);   // Sᵀ = K·Qᵀ | dPᵀ = V·dOᵀ
          commit_e(bar_s + buf);
        }
      } else {
        const uint32_t td = tmem + (is_v ? 256u : 320u), ta = tmem + (is_v ? 0u : 128u);
        if (n > 0) mbar_wait(bar_accfree, ph_item ^ 1u);             // the rows have read the previous item's dV / dK
        for (int i = 0; i < it.n_it; ++i) {
          const int gb = g + i, cur = gb & 1;
          mbar_wait(bar_mn + cur, (ph_a >> cur) & 1u); ph_a ^= 1u << cur;
          mbar_wait(bar_p + cur, (ph_b >> cur) & 1u); ph_b ^= 1u << cur;      // Pᵀ and dSᵀ are in tensor memory
          tc_fence_after();
          mma_tmn(td, ta + cur * 64u, aBmn + cur * kSmall, idesc_mn, i > 0);   // dV += Pᵀ·dO_i | dK += dSᵀ·Q_i
          commit_e(bar_o + cur);
        }
        commit_e(bar_done);
      }
      g += it.n_it;
      ++n;
      ph_item ^= 1u;
    }
  } else {
    const int quarter = warp & 3, half = (warp - 5) >> 2;
    const int r_in = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int shift = P.Sk - P.Sq;
    const float scale2 = P.scale * kLog2e, mask2 = P.mask_value * kLog2e;
    uint32_t ph_s = 0, ph_item = 0;
    int g = 0;
    for (int r = 0; r < rounds; ++r) {
      const DkvItem it = dkv_item(P, r);
      if (!it.valid) break;
      const int kv = it.kv0 + r_in;
      const bool kv_ok = kv < P.Sk;
      float *dk_row = P.g0 + (int64_t)it.b * P.g0_sb + (int64_t)it.h * P.g0_sh + (int64_t)kv * P.g0_ss + half * 32;
      float *dv_row = P.g1 + (int64_t)it.b * P.g1_sb + (int64_t)it.h * P.g1_sh + (int64_t)kv * P.g1_ss + half * 32;
      if (it.n_it == 0) {
        if (kv_ok) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            reinterpret_cast<uint4 *>(dv_row)[q] = make_uint4(0, 0, 0, 0);
            reinterpret_cast<uint4 *>(dk_row)[q] = make_uint4(0, 0, 0, 0);
          }
        }
        continue;
      }
      const uint8_t *mcol = P.mask ? P.mask + (int64_t)it.b * P.m_sb + (int64_t)it.h * P.m_sh + kv : nullptr;
      for (int i = 0; i < it.n_it; ++i) {
        const int gb = g + i, cur = gb & 1, sb = gb & 3;
        const int qc0 = (it.i_start + i) * kCols + half * 32;
        uint32_t mbits = 0;                                         // bit c: explicit mask at (query qc0 + c, key kv)
        if (mcol && kv_ok) {
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (qc0 + c < P.Sq && __ldg(mcol + (int64_t)(qc0 + c) * P.m_ss)) mbits |= 1u << c;
        }
        const bool fast = (qc0 + 32 <= P.Sq) && (it.kv0 + quarter * 32 + 32 <= P.Sk) &&
                          !(P.causal && it.kv0 + quarter * 32 + 31 > qc0 + shift) && !__any_sync(0xffffffffu, mbits != 0u);
        mbar_wait(bar_st + sb, (uint32_t)((gb >> 2) & 1));
        mbar_wait(bar_s + cur, (ph_s >> cur) & 1u); ph_s ^= 1u << cur;
        tc_fence_after();
        const uint32_t t_s = tmem + cur * 64u + lane_addr + half * 32, t_dp = tmem + 128u + cur * 64u + lane_addr + half * 32;
        uint32_t rs[32], rp[32];
        tmem_ld32(t_s, rs);
        tmem_ld32(t_dp, rp);
        tmem_ld_wait();
        const float4 *st = sStats + sb * kCols + half * 32;
        if (fast) {
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const float4 sq = st[c];                                // smem broadcast
            const float p = __fmul_rn(ex2(__fmaf_rn(__uint_as_float(rs[c]), scale2, -sq.x)), sq.y);
            rs[c] = __float_as_uint(p);
            rp[c] = __float_as_uint(__fmul_rn(p, __fsub_rn(__uint_as_float(rp[c]), sq.z)));
          }
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const int q = qc0 + c;
            const bool q_ok = q < P.Sq;
            const float4 sq = st[c];                                // smem broadcast; garbage past Sq is never used
            const bool mk = (P.causal && kv > q + shift) || ((mbits >> c) & 1u);
            const float t = mk ? __fsub_rn(mask2, sq.x) : __fmaf_rn(__uint_as_float(rs[c]), scale2, -sq.x);
            const float p = (q_ok && kv_ok) ? __fmul_rn(ex2(t), sq.y) : 0.0f;
            const float d = __fmul_rn(p, __fsub_rn(__uint_as_float(rp[c]), sq.z));
            rs[c] = __float_as_uint(p);
            rp[c] = __float_as_uint((mk || !(q_ok && kv_ok)) ? 0.0f : d);
          }
        }
        tmem_st32(t_s, rs);                                         // Pᵀ over Sᵀ, dSᵀ over dPᵀ, in place
        tmem_st32(t_dp, rp);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_p + cur);
      }
      {
        mbar_wait(bar_done, ph_item);                               // every dV / dK product of the item has retired
        tc_fence_after();
        uint32_t rv[32], rk[32];
        tmem_ld32(tmem + 256u + lane_addr + half * 32, rv);
        tmem_ld32(tmem + 320u + lane_addr + half * 32, rk);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_accfree);                    // the accumulators may be overwritten by the next item
        if (kv_ok) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            reinterpret_cast<uint4 *>(dv_row)[q] = make_uint4(rv[q * 4], rv[q * 4 + 1], rv[q * 4 + 2], rv[q * 4 + 3]);
            reinterpret_cast<float4 *>(dk_row)[q] =
                make_float4(__fmul_rn(__uint_as_float(rk[q * 4]), P.scale), __fmul_rn(__uint_as_float(rk[q * 4 + 1]), P.scale),
                            __fmul_rn(__uint_as_float(rk[q * 4 + 2]), P.scale), __fmul_rn(__uint_as_float(rk[q * 4 + 3]), P.scale));
          }
        }
      }
      g += it.n_it;
      ph_item ^= 1u;
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace fa

// ------------------------------------------------------------------------------------------ host
static int32_t fa_operand(const b200_tensor *t, bool mn_major, int64_t B, int64_t H, Operand &o) {
  o.ptr = t->ptr;
  o.es = 4;
  o.mn_major = mn_major;
  o.s_mn = mn_major ? t->strides[3] : t->strides[2];
  o.s_k = mn_major ? t->strides[2] : t->strides[3];
  o.s_b[0] = 0; o.s_b[1] = t->strides[0]; o.s_b[2] = t->strides[1];
  o.bsz[0] = 1; o.bsz[1] = (int32_t)B; o.bsz[2] = (int32_t)H;
  B200_REQUIRE(t->strides[3] == 1 && ((uintptr_t)t->ptr % 16) == 0, B200_ERR_UNSUPPORTED,
               "attention operands need a contiguous, 16-byte aligned head dim");
  const int64_t mn = mn_major ? t->shape[3] : t->shape[2], kk = mn_major ? t->shape[2] : t->shape[3];
  B200_REQUIRE(tma_ok(o, mn, kk), B200_ERR_UNSUPPORTED, "attention operand strides must be multiples of 4 elements");
  return B200_OK;
}
static int32_t fa_map(CUtensorMap *map, const b200_tensor *t, bool mn_major, int box_rows, int64_t B, int64_t H) {
  Operand o;
  int32_t st = fa_operand(t, mn_major, B, H, o);
  if (st != B200_OK) return st;
  return mn_major ? make_tmap(map, o, t->shape[3], t->shape[2]) : make_tmap(map, o, t->shape[2], t->shape[3], box_rows);
}
static bool fa_rows_ok(const b200_tensor *t) {
  return t->strides[3] == 1 && ((uintptr_t)t->ptr % 16) == 0 && t->strides[0] % 4 == 0 && t->strides[1] % 4 == 0 && t->strides[2] % 4 == 0;
}
struct FaMask {
  const uint8_t *ptr;
  int64_t sb, sh, ss;
};
static int32_t fa_mask(const b200_tensor *mask, int64_t B, int64_t H, int64_t Sq, int64_t Sk, FaMask &m) {
  m.ptr = nullptr;
  m.sb = m.sh = m.ss = 0;
  if (!mask) return B200_OK;
  B200_REQUIRE(mask->rank == 4 && (mask->dtype == B200_BOOL || mask->dtype == B200_U8) && mask->ptr, B200_ERR_UNSUPPORTED,
               "attention mask must be bool [B|1, H|1, Sq, Sk]");
  B200_REQUIRE((mask->shape[0] == B || mask->shape[0] == 1) && (mask->shape[1] == H || mask->shape[1] == 1) &&
                   mask->shape[2] == Sq && mask->shape[3] == Sk,
               B200_ERR_SHAPE, "attention mask is not broadcastable to [B, H, Sq, Sk]");
  B200_REQUIRE(mask->strides[3] == 1 && Sk % 4 == 0 && ((uintptr_t)mask->ptr % 4) == 0 && mask->strides[2] % 4 == 0 &&
                   mask->strides[0] % 4 == 0 && mask->strides[1] % 4 == 0,
               B200_ERR_UNSUPPORTED, "attention mask needs Sk %% 4 == 0 and 4-byte-multiple strides");
  m.ptr = reinterpret_cast<const uint8_t *>(mask->ptr);
  m.sb = mask->shape[0] == 1 ? 0 : mask->strides[0];
  m.sh = mask->shape[1] == 1 ? 0 : mask->strides[1];
  m.ss = mask->strides[2];
  return B200_OK;
}
static int32_t fa_stats_ok(const b200_tensor *stats, int64_t B, int64_t H, int64_t Sq) {
  B200_REQUIRE(stats && stats->ptr && stats->rank == 4 && stats->dtype == B200_F32 && stats->shape[0] == B && stats->shape[1] == H &&
                   stats->shape[2] == Sq && stats->shape[3] == 4 && is_contiguous(*stats) && ((uintptr_t)stats->ptr % 16) == 0,
               B200_ERR_SHAPE, "attention stats must be a contiguous, 16-byte aligned f32 [B, H, Sq, 4] tensor");
  return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" int32_t b200_launch_attention_flash(const b200_tensor *q, const b200_tensor *k, const b200_tensor *v,
                                               const b200_tensor *mask, double scale, double mask_value, int32_t is_causal,
                                               const b200_tensor *out, const b200_tensor *stats, b200_stream s) {
  B200_REQUIRE(q && k && v && out && stats, B200_ERR_INVALID, "null argument");
  for (const b200_tensor *t : {q, k, v, out})
    B200_REQUIRE(t->rank == 4 && t->dtype == B200_F32 && t->ptr, B200_ERR_UNSUPPORTED, "attention operands must be f32 [B, H, S, D]");
  const int64_t B = q->shape[0], H = q->shape[1], Sq = q->shape[2], D = q->shape[3], Sk = k->shape[2], Dv = v->shape[3];
  B200_REQUIRE(k->shape[0] == B && k->shape[1] == H && k->shape[3] == D && v->shape[0] == B && v->shape[1] == H &&
                   v->shape[2] == Sk && out->shape[0] == B && out->shape[1] == H && out->shape[2] == Sq && out->shape[3] == Dv,
               B200_ERR_SHAPE, "attention shape mismatch");
  B200_REQUIRE(D == fa::HD && Dv == fa::HD, B200_ERR_UNSUPPORTED,
               "the fused attention kernels are built for head dim 64 (got %lld / %lld); use the op chain", (long long)D, (long long)Dv);
  int32_t st = fa_stats_ok(stats, B, H, Sq);
  if (st != B200_OK) return st;
  if (B == 0 || H == 0 || Sq == 0) return B200_OK;
  B200_REQUIRE(Sk > 0 && Sq < (1ll << 30) && Sk < (1ll << 30), B200_ERR_UNSUPPORTED, "attention sequence lengths out of range");
  fa::FwdParams P;
  memset(&P, 0, sizeof(P));
  if ((st = fa_map(&P.tma_q, q, false, fa::kRows, B, H)) != B200_OK) return st;
  if ((st = fa_map(&P.tma_k, k, false, fa::kCols, B, H)) != B200_OK) return st;
  if ((st = fa_map(&P.tma_v, v, true, 0, B, H)) != B200_OK) return st;
  B200_REQUIRE(fa_rows_ok(out), B200_ERR_UNSUPPORTED, "attention output needs a contiguous head dim and 16-byte-multiple strides");
  P.out = reinterpret_cast<float *>(out->ptr);
  P.o_sb = out->strides[0]; P.o_sh = out->strides[1]; P.o_ss = out->strides[2];
  P.stats = reinterpret_cast<float4 *>(stats->ptr);
  FaMask fm;
  if ((st = fa_mask(mask, B, H, Sq, Sk, fm)) != B200_OK) return st;
  P.mask = fm.ptr; P.m_sb = fm.sb; P.m_sh = fm.sh; P.m_ss = fm.ss;
  P.B = (int32_t)B; P.H = (int32_t)H; P.Sq = (int32_t)Sq; P.Sk = (int32_t)Sk;
  P.causal = is_causal ? 1 : 0;
  P.q_blocks = (int32_t)((Sq + fa::kRows - 1) / fa::kRows);
  P.scale = (float)scale;
  P.mask_value = (float)mask_value;
  const int64_t ctas = B * H * P.q_blocks;
  B200_REQUIRE(ctas < (1ll << 31), B200_ERR_UNSUPPORTED, "attention grid too large");
  const size_t smem = 1024 + fa::kBig + 4 * fa::kSmall + 128;
  if ((st = ensure_dyn_smem(reinterpret_cast<const void *>(fa::flash_fwd_kernel), smem, true)) != B200_OK) return st;
  fa::flash_fwd_kernel<<<(unsigned)ctas, 192, smem, resolve_stream(s)>>>(P);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int32_t b200_launch_attention_flash_backward(const b200_tensor *d_out, const b200_tensor *q, const b200_tensor *k,
                                                        const b200_tensor *v, const b200_tensor *out, const b200_tensor *stats,
                                                        const b200_tensor *mask, double scale, double mask_value,
                                                        int32_t is_causal, const b200_tensor *dq, const b200_tensor *dk,
                                                        const b200_tensor *dv, b200_stream s) {
  B200_REQUIRE(d_out && q && k && v && out && stats && dq && dk && dv, B200_ERR_INVALID, "null argument");
  for (const b200_tensor *t : {d_out, q, k, v, out, dq, dk, dv})
    B200_REQUIRE(t->rank == 4 && t->dtype == B200_F32 && t->ptr, B200_ERR_UNSUPPORTED, "attention operands must be f32 rank-4");
  const int64_t B = q->shape[0], H = q->shape[1], Sq = q->shape[2], D = q->shape[3], Sk = k->shape[2];
  B200_REQUIRE(D == fa::HD, B200_ERR_UNSUPPORTED, "the fused attention kernels are built for head dim 64; use the op chain");
  auto same = [&](const b200_tensor *t, int64_t s2) {
    return t->shape[0] == B && t->shape[1] == H && t->shape[2] == s2 && t->shape[3] == D;
  };
  B200_REQUIRE(same(d_out, Sq) && same(out, Sq) && same(dq, Sq) && same(k, Sk) && same(v, Sk) && same(dk, Sk) && same(dv, Sk),
               B200_ERR_SHAPE, "attention backward shape mismatch");
  int32_t st = fa_stats_ok(stats, B, H, Sq);
  if (st != B200_OK) return st;
  if (B == 0 || H == 0 || Sq == 0 || Sk == 0) return B200_OK;
  B200_REQUIRE(Sq < (1ll << 30) && Sk < (1ll << 30), B200_ERR_UNSUPPORTED, "attention sequence lengths out of range");
  B200_REQUIRE(fa_rows_ok(out) && fa_rows_ok(d_out) && fa_rows_ok(dq) && fa_rows_ok(dk) && fa_rows_ok(dv), B200_ERR_UNSUPPORTED,
               "attention rows need a contiguous head dim and 16-byte-multiple strides");
  FaMask fm;
  if ((st = fa_mask(mask, B, H, Sq, Sk, fm)) != B200_OK) return st;
  fa::BwdParams P;
  memset(&P, 0, sizeof(P));
  P.o_fwd = reinterpret_cast<const float *>(out->ptr);
  P.of_sb = out->strides[0]; P.of_sh = out->strides[1]; P.of_ss = out->strides[2];
  P.d_out = reinterpret_cast<const float *>(d_out->ptr);
  P.do_sb = d_out->strides[0]; P.do_sh = d_out->strides[1]; P.do_ss = d_out->strides[2];
  P.stats = reinterpret_cast<float4 *>(stats->ptr);
  P.mask = fm.ptr; P.m_sb = fm.sb; P.m_sh = fm.sh; P.m_ss = fm.ss;
  P.B = (int32_t)B; P.H = (int32_t)H; P.Sq = (int32_t)Sq; P.Sk = (int32_t)Sk;
  P.causal = is_causal ? 1 : 0;
  P.scale = (float)scale;
  P.mask_value = (float)mask_value;
  cudaStream_t stream = resolve_stream(s);

  // ---- kernel 1: dQ (and delta into stats.z)
  if ((st = fa_map(&P.tma_q, q, false, fa::kRows, B, H)) != B200_OK) return st;
  if ((st = fa_map(&P.tma_do, d_out, false, fa::kRows, B, H)) != B200_OK) return st;
  if ((st = fa_map(&P.tma_k, k, false, fa::kCols, B, H)) != B200_OK) return st;
  if ((st = fa_map(&P.tma_v, v, false, fa::kCols, B, H)) != B200_OK) return st;
  if ((st = fa_map(&P.tma_mn0, k, true, 0, B, H)) != B200_OK) return st;
  if ((st = fa_map(&P.tma_mn1, out, false, fa::kRows, B, H)) != B200_OK) return st;   // forward output tile, for delta
  P.g0 = reinterpret_cast<float *>(dq->ptr);
  P.g0_sb = dq->strides[0]; P.g0_sh = dq->strides[1]; P.g0_ss = dq->strides[2];
  P.blocks = (int32_t)((Sq + fa::kRows - 1) / fa::kRows);
  {
    const int64_t ctas = B * H * P.blocks;
    B200_REQUIRE(ctas < (1ll << 31), B200_ERR_UNSUPPORTED, "attention grid too large");
    constexpr int kParts = 2;   // 8 row warps (16 measured 3 % slower: the tensor pipe and the tile traffic bound this kernel, not the row math)
    const size_t smem = 1024 + 3 * fa::kBig + 6 * fa::kSmall + kParts * fa::kRows * 4 + 128;
    if ((st = ensure_dyn_smem(reinterpret_cast<const void *>(fa::flash_bwd_dq_kernel<kParts>), smem)) != B200_OK) return st;
    static const bool persist = [] { const char *e = std::getenv("B200_FA_PERSIST"); return !(e && e[0] == '0'); }();
    // where it was measured to win: causal masks (balance) and short key loops (per-item cost dominates)
    const bool skip_blocks = P.causal && !(Sq > Sk && P.mask_value > -INFINITY);
    if (persist && (skip_blocks || (Sk + fa::kCols - 1) / fa::kCols <= 8)) {
      const size_t psmem = smem + 64;
      if ((st = ensure_dyn_smem(reinterpret_cast<const void *>(fa::flash_bwd_dq_persist_kernel<kParts>), psmem)) != B200_OK) return st;
      const unsigned grid = (unsigned)std::min<int64_t>(ctas, sm_count());
      fa::flash_bwd_dq_persist_kernel<kParts><<<grid, 128 + 128 * kParts, psmem, stream>>>(P);
    } else
    fa::flash_bwd_dq_kernel<kParts><<<(unsigned)ctas, 128 + 128 * kParts, smem, stream>>>(P);
    B200_LAUNCH_CHECK();
  }
  // ---- kernel 2: dK, dV
  if ((st = fa_map(&P.tma_k, k, false, fa::kRows, B, H)) != B200_OK) return st;
  if ((st = fa_map(&P.tma_v, v, false, fa::kRows, B, H)) != B200_OK) return st;
  if ((st = fa_map(&P.tma_q, q, false, fa::kCols, B, H)) != B200_OK) return st;
  if ((st = fa_map(&P.tma_do, d_out, false, fa::kCols, B, H)) != B200_OK) return st;
  if ((st = fa_map(&P.tma_mn0, q, true, 0, B, H)) != B200_OK) return st;
  if ((st = fa_map(&P.tma_mn1, d_out, true, 0, B, H)) != B200_OK) return st;
  P.g0 = reinterpret_cast<float *>(dk->ptr);
  P.g0_sb = dk->strides[0]; P.g0_sh = dk->strides[1]; P.g0_ss = dk->strides[2];
  P.g1 = reinterpret_cast<float *>(dv->ptr);
  P.g1_sb = dv->strides[0]; P.g1_sh = dv->strides[1]; P.g1_ss = dv->strides[2];
  P.blocks = (int32_t)((Sk + fa::kRows - 1) / fa::kRows);
  {
    const int64_t ctas = B * H * P.blocks;
    B200_REQUIRE(ctas < (1ll << 31), B200_ERR_UNSUPPORTED, "attention grid too large");
    const size_t smem = 1024 + 2 * fa::kBig + 8 * fa::kSmall + 4 * fa::kCols * 16 + 256;
    if ((st = ensure_dyn_smem(reinterpret_cast<const void *>(fa::flash_bwd_dkv_kernel), smem)) != B200_OK) return st;
    static const bool persist = [] { const char *e = std::getenv("B200_FA_PERSIST"); return !(e && e[0] == '0'); }();
    // measured: [64,8,256,64] (four query blocks per item) backward 155 -> 141 us; on [8,16,1024,64] causal the per-item
    // kernel with the heaviest-first launch order is 9 us faster — short query loops only
    if (persist && (Sq + fa::kCols - 1) / fa::kCols <= 8) {
      const size_t psmem = smem + 64;
      if ((st = ensure_dyn_smem(reinterpret_cast<const void *>(fa::flash_bwd_dkv_persist_kernel), psmem)) != B200_OK) return st;
      const unsigned grid = (unsigned)std::min<int64_t>(ctas, sm_count());
      fa::flash_bwd_dkv_persist_kernel<<<grid, fa::kDkvThreads, psmem, stream>>>(P);
    } else
    fa::flash_bwd_dkv_kernel<<<(unsigned)ctas, fa::kDkvThreads, smem, stream>>>(P);
    B200_LAUNCH_CHECK();
  }
  return B200_OK;
}
