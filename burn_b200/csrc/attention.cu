// Fused attention forward: out = softmax(q·kᵀ·scale [masked → mask_value]) · v in one kernel
// (SURVEY.md §8(f) row 2).  Replaces ModuleOps::attention
// (crates/burn-backend/src/backend/ops/modules/base.rs:822-830) whose semantics are pinned by
// attention_fallback (crates/burn-backend/src/backend/ops/modules/attention.rs:15-90): scores scaled,
// bool mask / causal mask filled, NaN-safe softmax (row max clamped to the most negative finite
// value, row sum to the smallest normal), context = weights · v.  With mask_value = -1e9 and the
// weights returned it is also the forward of burn-nn's MultiHeadAttention core
// (crates/burn-nn/src/modules/attention/mha.rs:253-311), whose backward needs the weights.
//
// Design (tcgen05, head dim 64, f32 storage, tf32 tensor-core math, f32 softmax):
//   (softmax in the base-2 domain with MUFU.EX2 and one reciprocal per row — see the kernel)
//   one CTA per (batch, head, 128 query rows); 192 threads = TMA/MMA issuer warp, TMEM-allocator warp,
//   4 softmax warps (one thread per query row = one TMEM lane).  Two passes over the 64-wide KV blocks:
//     pass A  S = Q·K_jᵀ (tcgen05.mma → TMEM) → running row max m and row sum l
//     pass B  S again → P = exp(S - m) / l, exactly the reference's formula → written to smem as the
//             K-major A operand (and to the optional weights tensor) → O += P·V_j accumulated in TMEM
//   so the [B,H,Sq,Sk] score tensor never goes to HBM (only the weights, once, when asked for), and O
//   needs no rescaling.  Fully masked KV blocks of a causal mask are skipped (their weights are exact
//   zeros).  smem 96 KB + 128 TMEM columns per CTA: two CTAs per SM overlap each other's TMA / MMA /
//   softmax phases.  Roofline: HBM when weights are written (4·Sq·Sk bytes per head), tensor pipe
//   otherwise; FLOPs = 6·Sq·Sk·D per head (QKᵀ twice + PV).
#include "tcgen05.cuh"

namespace b200 {
namespace attn {

using namespace mm;

constexpr int BQ = 128, BKV = 64, HD = 64;
constexpr int kThreads = 192;
constexpr uint32_t kQBytes = BQ * HD * 4, kKBytes = BKV * HD * 4, kVBytes = BKV * HD * 4, kPBytes = BQ * BKV * 4;
constexpr uint32_t kMBytes = BQ * BKV;   // one mask tile (bytes)

struct Params {
  CUtensorMap tma_q, tma_k, tma_v, tma_w;   // tma_w: weights [Sk, Sq, H, B] boxes of 32 x 128, 128B swizzle
  float *out;
  int64_t o_sb, o_sh, o_ss;
  float *w;                      // optional weights [B,H,Sq,Sk]
  int64_t w_sb, w_sh, w_ss;
  const uint8_t *mask;           // optional bool mask, nonzero = masked
  int64_t m_sb, m_sh, m_ss;
  CUtensorMap tma_m;             // mask tiles [64 columns x 128 rows] through TMA when its strides allow (mask_tma)
  int32_t mask_tma, m_bcast_b, m_bcast_h;
  int32_t B, H, Sq, Sk;
  int32_t causal, q_blocks;
  float scale, mask_value;
  // backward (attention_bwd_dq_kernel): tma_q = dO, tma_k = V (K-major), tma_v = K (MN-major), tma_p loads the
  // weights tile, tma_w stores dS, out = dQ; o_fwd / d_out are read per row for delta = sum(dO * O)
  CUtensorMap tma_p;
  const float *o_fwd, *d_out;
  int64_t of_sb, of_sh, of_ss, do_sb, do_sh, do_ss;
};

// MASK_TMA: the bool mask arrives as TMA-loaded [128 x 64 B] tiles (kept out of the unmasked / causal
// instantiation, whose row loop is the hot path)
template <bool MASK_TMA>
__global__ void __launch_bounds__(kThreads, 2) attention_fwd_kernel(const __grid_constant__ Params P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *sQ = smem, *sK = sQ + kQBytes, *sV = sK + kKBytes, *sP = sV + kVBytes;
  uint8_t *sM = sP + kPBytes;                      // mask tile [128 rows][64 bytes]
  uint64_t *bars = reinterpret_cast<uint64_t *>(sM + (MASK_TMA ? kMBytes : 0));
  uint64_t *bar_q = bars, *bar_k = bars + 1, *bar_v = bars + 2, *bar_s = bars + 3, *bar_p = bars + 4, *bar_o = bars + 5,
           *bar_m = bars + 6;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 7);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // heavy (late, under a causal mask) query blocks first
  const int work = blockIdx.x;
  const int qb = P.q_blocks - 1 - (work % P.q_blocks);
  const int bh = work / P.q_blocks, h = bh % P.H, b = bh / P.H;
  const int q0 = qb * BQ;
  const int nkv_all = (P.Sk + BKV - 1) / BKV;
  int nkv = nkv_all;
  if (P.causal) {
    const int last_col = min(P.Sk - 1, q0 + BQ - 1 + (P.Sk - P.Sq));   // last visible column of the block's last row
    nkv = last_col < 0 ? 0 : min(nkv_all, last_col / BKV + 1);
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tma_q);
    tma_prefetch_desc(&P.tma_k);
    tma_prefetch_desc(&P.tma_v);
    if (P.w) tma_prefetch_desc(&P.tma_w);
    mbar_init(bar_q, 1);
    mbar_init(bar_k, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 4);
    mbar_init(bar_o, 2);      // tcgen05.commit of P·V + the issuer once the weights store has read sP
    mbar_init(bar_m, 1);
    if constexpr (MASK_TMA) tma_prefetch_desc(&P.tma_m);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = *tmem_slot, tmem_o = tmem_s + 64;

  if (warp == 0) {
    if (lane == 0 && nkv > 0) {
      // ===================== TMA + MMA issuer =====================
      uint32_t idesc = 0;
      idesc |= 1u << 4;                     // D = f32
      idesc |= 2u << 7;                     // A = tf32
      idesc |= 2u << 10;                    // B = tf32
      idesc |= (uint32_t)(BKV >> 3) << 17;  // N = 64
      idesc |= (uint32_t)(BQ >> 4) << 24;   // M = 128
      const uint32_t idesc_pv = idesc | (1u << 16);   // B (V) is MN-major
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP);
      mbar_expect_tx(bar_q, kQBytes);
      tma_load_5d(sQ, &P.tma_q, bar_q, 0, q0, h, b, 0);
      tma_load_5d(sQ + kQBytes / 2, &P.tma_q, bar_q, 32, q0, h, b, 0);
      mbar_wait(bar_q, 0);
      uint32_t ph_k = 0, ph_v = 0, ph_p = 0, ph_o = 0;
      const int mb = P.m_bcast_b ? 0 : b, mh = P.m_bcast_h ? 0 : h;
      auto load_k = [&](int j) {
        if constexpr (MASK_TMA) {   // the mask tile travels with K: its buffer is free once S(j-1) was consumed
          mbar_expect_tx(bar_m, kMBytes);
          tma_load_5d(sM, &P.tma_m, bar_m, j * BKV, q0, mh, mb, 0);
        }
        mbar_expect_tx(bar_k, kKBytes);
        tma_load_5d(sK, &P.tma_k, bar_k, 0, j * BKV, h, b, 0);
        tma_load_5d(sK + kKBytes / 2, &P.tma_k, bar_k, 32, j * BKV, h, b, 0);
      };
      auto mma_qk = [&]() {
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < HD / 8; ++k) {
          const uint64_t da = make_desc(aQ + (k >> 2) * (kQBytes / 2) + (k & 3) * 32, 16, 1024);
          const uint64_t db = make_desc(aK + (k >> 2) * (kKBytes / 2) + (k & 3) * 32, 16, 1024);
          umma<false>(tmem_s, da, db, idesc, k ? 1u : 0u);
        }
        umma_commit(bar_s);
      };
      // ---- pass A: row statistics
      for (int j = 0; j < nkv; ++j) {
        if (j > 0) { mbar_wait(bar_p, ph_p); ph_p ^= 1; }       // S(j-1) consumed → S and sK are free
        load_k(j);
        mbar_wait(bar_k, ph_k); ph_k ^= 1;
        mma_qk();
      }
      mbar_wait(bar_p, ph_p); ph_p ^= 1;
      // ---- pass B: weights and context
      for (int j = 0; j < nkv; ++j) {
        load_k(j);
        if (j > 0) { mbar_wait(bar_o, ph_o); ph_o ^= 1; }       // P·V(j-1) retired → sV and sP are free
        mbar_expect_tx(bar_v, kVBytes);
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int g = 0; g < 2; ++g)
            tma_load_5d(sV + kb * 8192 + g * 4096, &P.tma_v, bar_v, g * 32, j * BKV + kb * 32, h, b, 0);
        mbar_wait(bar_k, ph_k); ph_k ^= 1;
        mma_qk();
        mbar_wait(bar_v, ph_v); ph_v ^= 1;
        mbar_wait(bar_p, ph_p); ph_p ^= 1;                      // P(j) is in smem, S(j) consumed
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < BKV / 8; ++k) {
          const uint64_t da = make_desc(aP + (k >> 2) * (kPBytes / 2) + (k & 3) * 32, 16, 1024);
          const uint64_t db = make_desc(aV + (k >> 2) * 8192 + (k & 3) * 1024, 4096, 512, 1);
          umma<false>(tmem_o, da, db, idesc_pv, (j | k) ? 1u : 0u);
        }
        umma_commit(bar_o);
        if (P.w) {
          // sP is exactly two [128 rows x 32 columns] 128B-swizzled TMA boxes: the weights leave as
          // full-line bulk stores (rows / columns past Sq / Sk are clipped by the tensor map)
          tma_store_5d(&P.tma_w, sP, j * BKV, q0, h, b, 0);
          tma_store_5d(&P.tma_w, sP + kPBytes / 2, j * BKV + 32, q0, h, b, 0);
          bulk_commit();
          bulk_wait_read<0>();
        }
        mbar_arrive(bar_o);
      }
      if (P.w) bulk_wait_all();
    }
  } else if (warp >= 2) {
    // ===================== softmax warps: thread = query row = TMEM lane =====================
    const int quarter = warp & 3;
    const int r_in = quarter * 32 + lane, row = q0 + r_in;
    const bool row_ok = row < P.Sq;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int causal_limit = row + (P.Sk - P.Sq);                 // columns beyond are masked when causal
    const uint8_t *mrow = P.mask ? P.mask + (int64_t)b * P.m_sb + (int64_t)h * P.m_sh + (int64_t)row * P.m_ss : nullptr;
    uint32_t ph_s = 0, ph_o = 0, ph_m = 0;
    float m = -INFINITY, l = 0.0f;

    // Softmax runs in the base-2 domain: s2 = score·scale·log2(e), p = 2^(s2 - m2) / l — one FMUL and one
    // MUFU.EX2 per element instead of the ~30-instruction expf + IEEE division (the kernel is issue-bound
    // on exactly this math: 128x64 elements per KV block on 4 warps).  ex2.approx is good to 2 ulp and the
    // single reciprocal per row to 1 ulp — both far inside the tf32 products' 2^-11.
    const float kLog2e = 1.4426950408889634f;
    const float scale2 = P.scale * kLog2e, mask2 = P.mask_value * kLog2e;   // -inf stays -inf
    auto ex2 = [](float x) {
      float y;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
      return y;
    };
    const bool mask16 = mrow && (((uintptr_t)P.mask | (uintptr_t)P.m_sb | (uintptr_t)P.m_sh | (uintptr_t)P.m_ss) % 16 == 0) &&
                        P.Sk % 16 == 0;
    // scaled, masked scores of KV block j for this row; columns past Sk are -inf (excluded)
    auto scores = [&](int j, float (&s)[BKV]) {
      uint32_t r0[32], r1[32];
      tmem_ld32(tmem_s + lane_addr, r0);
      tmem_ld32(tmem_s + lane_addr + 32, r1);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        s[c] = __fmul_rn(__uint_as_float(r0[c]), scale2);
        s[c + 32] = __fmul_rn(__uint_as_float(r1[c]), scale2);
      }
      const int col0 = j * BKV;
      if constexpr (MASK_TMA) {
        mbar_wait(bar_m, ph_m); ph_m ^= 1;
        const uint4 *mt = reinterpret_cast<const uint4 *>(sM + r_in * BKV);
#pragma unroll
        for (int q = 0; q < BKV / 16; ++q) {
          const uint4 mw = mt[q];
          const uint32_t w4[4] = {mw.x, mw.y, mw.z, mw.w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            if (w4[t] & 0xFFu) s[q * 16 + t * 4] = mask2;
            if (w4[t] & 0xFF00u) s[q * 16 + t * 4 + 1] = mask2;
            if (w4[t] & 0xFF0000u) s[q * 16 + t * 4 + 2] = mask2;
            if (w4[t] & 0xFF000000u) s[q * 16 + t * 4 + 3] = mask2;
          }
        }
      } else if (mrow && row_ok) {
        if (mask16) {
#pragma unroll
          for (int q = 0; q < BKV / 16; ++q) {
            if (col0 + q * 16 < P.Sk) {
              const uint4 mw = __ldg(reinterpret_cast<const uint4 *>(mrow + col0) + q);
              const uint32_t w4[4] = {mw.x, mw.y, mw.z, mw.w};
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                if (w4[t] & 0xFFu) s[q * 16 + t * 4] = mask2;
                if (w4[t] & 0xFF00u) s[q * 16 + t * 4 + 1] = mask2;
                if (w4[t] & 0xFF0000u) s[q * 16 + t * 4 + 2] = mask2;
                if (w4[t] & 0xFF000000u) s[q * 16 + t * 4 + 3] = mask2;
              }
            }
          }
        } else {
#pragma unroll
          for (int q = 0; q < BKV / 4; ++q) {
            if (col0 + q * 4 < P.Sk) {
              const uint32_t mw = __ldg(reinterpret_cast<const uint32_t *>(mrow + col0) + q);
              if (mw & 0xFFu) s[q * 4] = mask2;
              if (mw & 0xFF00u) s[q * 4 + 1] = mask2;
              if (mw & 0xFF0000u) s[q * 4 + 2] = mask2;
              if (mw & 0xFF000000u) s[q * 4 + 3] = mask2;
            }
          }
        }
      }
      // only the blocks that cross the causal diagonal / the end of the keys need per-element checks
      if (P.causal && col0 + BKV - 1 > q0 + (P.Sk - P.Sq)) {
#pragma unroll
        for (int c = 0; c < BKV; ++c)
          if (col0 + c > causal_limit) s[c] = mask2;
      }
      if (col0 + BKV > P.Sk) {
#pragma unroll
        for (int c = 0; c < BKV; ++c)
          if (col0 + c >= P.Sk) s[c] = -INFINITY;
      }
    };

    // ---- pass A
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(bar_s, ph_s); ph_s ^= 1;
      tc_fence_after();
      float s[BKV];
      scores(j, s);
      float bm = s[0];
#pragma unroll
      for (int c = 1; c < BKV; ++c) bm = fmaxf(bm, s[c]);
      const float m_new = fmaxf(m, bm);
      const float m_use = fmaxf(m_new, -3.402823466e+38f);
      float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll
      for (int c = 0; c < BKV; c += 2) {
        acc0 += ex2(s[c] - m_use);
        acc1 += ex2(s[c + 1] - m_use);
      }
      l = l * ex2(m - m_use) + (acc0 + acc1);
      m = m_new;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
    }
    const float m_fin = fmaxf(m, -3.402823466e+38f);            // finfo.min clamp (attention.rs:70-72)
    const float l_fin = fmaxf(l, 1.175494351e-38f);             // min_positive clamp (attention.rs:75-76)
    const float rinv = __frcp_rn(l_fin);

    // ---- pass B
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(bar_s, ph_s); ph_s ^= 1;
      tc_fence_after();
      float s[BKV];
      scores(j, s);
#pragma unroll
      for (int c = 0; c < BKV; ++c) s[c] = ex2(s[c] - m_fin) * rinv;
      if (j > 0) { mbar_wait(bar_o, ph_o); ph_o ^= 1; }         // the previous P·V no longer reads sP
      // K-major A operand, 128-byte swizzle: 16-byte chunk q of row r lives at chunk q ^ (r & 7)
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float4 *>(sP + kb * (kPBytes / 2) + r_in * 128 + ((q ^ (r_in & 7)) << 4)) =
              make_float4(s[kb * 32 + q * 4], s[kb * 32 + q * 4 + 1], s[kb * 32 + q * 4 + 2], s[kb * 32 + q * 4 + 3]);
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
    }
    // weights of the skipped (fully masked) blocks are exact zeros: each warp walks its 32 rows and
    // writes 512 contiguous bytes per instruction
    if (P.w && nkv * BKV < P.Sk) {
      const int c_begin = nkv * BKV;
      for (int rr = 0; rr < 32; ++rr) {
        const int zr = q0 + quarter * 32 + rr;
        if (zr >= P.Sq) break;
        float *zrow = P.w + (int64_t)b * P.w_sb + (int64_t)h * P.w_sh + (int64_t)zr * P.w_ss;
        for (int c = c_begin + lane * 4; c < P.Sk; c += 128) __stcs(reinterpret_cast<float4 *>(zrow + c), make_float4(0.f, 0.f, 0.f, 0.f));
      }
    }
    // ---- context row
    float *orow = P.out + (int64_t)b * P.o_sb + (int64_t)h * P.o_sh + (int64_t)row * P.o_ss;
    if (nkv > 0) {
      mbar_wait(bar_o, ph_o); ph_o ^= 1;
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld32(tmem_o + lane_addr, o0);
      tmem_ld32(tmem_o + lane_addr + 32, o1);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          reinterpret_cast<uint4 *>(orow)[q] = make_uint4(o0[q * 4], o0[q * 4 + 1], o0[q * 4 + 2], o0[q * 4 + 3]);
          reinterpret_cast<uint4 *>(orow)[q + 8] = make_uint4(o1[q * 4], o1[q * 4 + 1], o1[q * 4 + 2], o1[q * 4 + 3]);
        }
      }
    } else if (row_ok) {
#pragma unroll
      for (int q = 0; q < 16; ++q) reinterpret_cast<uint4 *>(orow)[q] = make_uint4(0, 0, 0, 0);
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_s, 128);
  }
}

// Backward, first half: per 128 query rows, dP = dO·Vᵀ → dS = P ∘ (dP − delta)·scale → dQ = dS·K, with
// delta = rowsum(dO ∘ O) (= rowsum(dP ∘ P), so dP never needs a second pass).  Same pipeline as the
// forward's pass B with (Q, K, V) → (dO, V, K): dP in TMEM, the weights tile TMA-loaded into the sP
// buffer, turned into dS in place by the row threads, consumed from there by the dQ MMA and TMA-stored
// as the dS tensor the dK = dSᵀ·Q product reads.  Masked positions have P = 0, hence dS = 0: no mask
// is needed; causally skipped blocks are zero-filled.  [B,H,Sq,Sk] traffic: read P once, write dS once
// (the unfused chain: write dP, read dP and P, write dS, read dS).
__global__ void __launch_bounds__(kThreads, 2) attention_bwd_dq_kernel(const __grid_constant__ Params P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *sQ = smem, *sK = sQ + kQBytes, *sV = sK + kKBytes, *sP = sV + kVBytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(sP + kPBytes);
  uint64_t *bar_q = bars, *bar_k = bars + 1, *bar_v = bars + 2, *bar_s = bars + 3, *bar_p = bars + 4, *bar_o = bars + 5,
           *bar_pl = bars + 6;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 7);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int work = blockIdx.x;
  const int qb = P.q_blocks - 1 - (work % P.q_blocks);
  const int bh = work / P.q_blocks, h = bh % P.H, b = bh / P.H;
  const int q0 = qb * BQ;
  const int nkv_all = (P.Sk + BKV - 1) / BKV;
  int nkv = nkv_all;
  if (P.causal) {
    const int last_col = min(P.Sk - 1, q0 + BQ - 1 + (P.Sk - P.Sq));
    nkv = last_col < 0 ? 0 : min(nkv_all, last_col / BKV + 1);
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&P.tma_q);
    tma_prefetch_desc(&P.tma_k);
    tma_prefetch_desc(&P.tma_v);
    tma_prefetch_desc(&P.tma_p);
    tma_prefetch_desc(&P.tma_w);
    mbar_init(bar_q, 1);
    mbar_init(bar_k, 1);
    mbar_init(bar_v, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 4);
    mbar_init(bar_o, 2);
    mbar_init(bar_pl, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = *tmem_slot, tmem_o = tmem_s + 64;

  if (warp == 0) {
    if (lane == 0 && nkv > 0) {
      uint32_t idesc = 0;
      idesc |= 1u << 4;
      idesc |= 2u << 7;
      idesc |= 2u << 10;
      idesc |= (uint32_t)(BKV >> 3) << 17;
      idesc |= (uint32_t)(BQ >> 4) << 24;
      const uint32_t idesc_mn = idesc | (1u << 16);
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP);
      mbar_expect_tx(bar_q, kQBytes);
      tma_load_5d(sQ, &P.tma_q, bar_q, 0, q0, h, b, 0);
      tma_load_5d(sQ + kQBytes / 2, &P.tma_q, bar_q, 32, q0, h, b, 0);
      mbar_wait(bar_q, 0);
      uint32_t ph_k = 0, ph_v = 0, ph_p = 0, ph_o = 0;
      for (int j = 0; j < nkv; ++j) {
        mbar_expect_tx(bar_k, kKBytes);                             // V_j as the K-major B operand of dP
        tma_load_5d(sK, &P.tma_k, bar_k, 0, j * BKV, h, b, 0);
        tma_load_5d(sK + kKBytes / 2, &P.tma_k, bar_k, 32, j * BKV, h, b, 0);
        if (j > 0) { mbar_wait(bar_o, ph_o); ph_o ^= 1; }           // dQ MMA and dS store of j-1 are done with sV / sP
        mbar_expect_tx(bar_v, kVBytes);                             // K_j as the MN-major B operand of dQ
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int g = 0; g < 2; ++g)
            tma_load_5d(sV + kb * 8192 + g * 4096, &P.tma_v, bar_v, g * 32, j * BKV + kb * 32, h, b, 0);
        mbar_expect_tx(bar_pl, kPBytes);                            // the weights tile, in the A-operand layout
        tma_load_5d(sP, &P.tma_p, bar_pl, j * BKV, q0, h, b, 0);
        tma_load_5d(sP + kPBytes / 2, &P.tma_p, bar_pl, j * BKV + 32, q0, h, b, 0);
        mbar_wait(bar_k, ph_k); ph_k ^= 1;
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < HD / 8; ++k) {
          const uint64_t da = make_desc(aQ + (k >> 2) * (kQBytes / 2) + (k & 3) * 32, 16, 1024);
          const uint64_t db = make_desc(aK + (k >> 2) * (kKBytes / 2) + (k & 3) * 32, 16, 1024);
          umma<false>(tmem_s, da, db, idesc, k ? 1u : 0u);
        }
        umma_commit(bar_s);
        mbar_wait(bar_v, ph_v); ph_v ^= 1;
        mbar_wait(bar_p, ph_p); ph_p ^= 1;                          // dS(j) is in sP, dP(j) consumed
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < BKV / 8; ++k) {
          const uint64_t da = make_desc(aP + (k >> 2) * (kPBytes / 2) + (k & 3) * 32, 16, 1024);
          const uint64_t db = make_desc(aV + (k >> 2) * 8192 + (k & 3) * 1024, 4096, 512, 1);
          umma<false>(tmem_o, da, db, idesc_mn, (j | k) ? 1u : 0u);
        }
        umma_commit(bar_o);
        tma_store_5d(&P.tma_w, sP, j * BKV, q0, h, b, 0);
        tma_store_5d(&P.tma_w, sP + kPBytes / 2, j * BKV + 32, q0, h, b, 0);
        bulk_commit();
        bulk_wait_read<0>();
        mbar_arrive(bar_o);
      }
      bulk_wait_all();
    }
  } else if (warp >= 2) {
    const int quarter = warp & 3;
    const int r_in = quarter * 32 + lane, row = q0 + r_in;
    const bool row_ok = row < P.Sq;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    // delta = sum_d dO[row, d] * O[row, d]
    float delta = 0.0f;
    if (row_ok) {
      const float4 *orow = reinterpret_cast<const float4 *>(P.o_fwd + (int64_t)b * P.of_sb + (int64_t)h * P.of_sh + (int64_t)row * P.of_ss);
      const float4 *grow = reinterpret_cast<const float4 *>(P.d_out + (int64_t)b * P.do_sb + (int64_t)h * P.do_sh + (int64_t)row * P.do_ss);
#pragma unroll
      for (int q = 0; q < HD / 4; ++q) {
        const float4 o = __ldg(orow + q), g = __ldg(grow + q);
        delta = __fadd_rn(delta, __fadd_rn(__fadd_rn(__fmul_rn(o.x, g.x), __fmul_rn(o.y, g.y)),
                                           __fadd_rn(__fmul_rn(o.z, g.z), __fmul_rn(o.w, g.w))));
      }
    }
    uint32_t ph_s = 0, ph_o = 0, ph_pl = 0;
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(bar_s, ph_s); ph_s ^= 1;
      tc_fence_after();
      uint32_t r0[32], r1[32];
      tmem_ld32(tmem_s + lane_addr, r0);
      tmem_ld32(tmem_s + lane_addr + 32, r1);
      tmem_ld_wait();
      mbar_wait(bar_pl, ph_pl); ph_pl ^= 1;
      // the tile only lands after bar_o(j-1); waiting it here (satisfied already) keeps this thread's
      // phase tracking of bar_o in lock step, so the final wait below is for the LAST completion
      if (j > 0) { mbar_wait(bar_o, ph_o); ph_o ^= 1; }
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 *cell = reinterpret_cast<float4 *>(sP + kb * (kPBytes / 2) + r_in * 128 + ((q ^ (r_in & 7)) << 4));
          const float4 p = *cell;
          const uint32_t *dp = kb ? r1 : r0;
          float4 d;
          d.x = __fmul_rn(__fmul_rn(p.x, __fsub_rn(__uint_as_float(dp[q * 4]), delta)), P.scale);
          d.y = __fmul_rn(__fmul_rn(p.y, __fsub_rn(__uint_as_float(dp[q * 4 + 1]), delta)), P.scale);
          d.z = __fmul_rn(__fmul_rn(p.z, __fsub_rn(__uint_as_float(dp[q * 4 + 2]), delta)), P.scale);
          d.w = __fmul_rn(__fmul_rn(p.w, __fsub_rn(__uint_as_float(dp[q * 4 + 3]), delta)), P.scale);
          *cell = d;
        }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p);
    }
    if (nkv * BKV < P.Sk) {   // dS of the causally skipped blocks
      const int c_begin = nkv * BKV;
      for (int rr = 0; rr < 32; ++rr) {
        const int zr = q0 + quarter * 32 + rr;
        if (zr >= P.Sq) break;
        float *zrow = P.w + (int64_t)b * P.w_sb + (int64_t)h * P.w_sh + (int64_t)zr * P.w_ss;
        for (int c = c_begin + lane * 4; c < P.Sk; c += 128) __stcs(reinterpret_cast<float4 *>(zrow + c), make_float4(0.f, 0.f, 0.f, 0.f));
      }
    }
    float *qrow = P.out + (int64_t)b * P.o_sb + (int64_t)h * P.o_sh + (int64_t)row * P.o_ss;
    if (nkv > 0) {
      mbar_wait(bar_o, ph_o); ph_o ^= 1;
      tc_fence_after();
      uint32_t o0[32], o1[32];
      tmem_ld32(tmem_o + lane_addr, o0);
      tmem_ld32(tmem_o + lane_addr + 32, o1);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          reinterpret_cast<uint4 *>(qrow)[q] = make_uint4(o0[q * 4], o0[q * 4 + 1], o0[q * 4 + 2], o0[q * 4 + 3]);
          reinterpret_cast<uint4 *>(qrow)[q + 8] = make_uint4(o1[q * 4], o1[q * 4 + 1], o1[q * 4 + 2], o1[q * 4 + 3]);
        }
      }
    } else if (row_ok) {
#pragma unroll
      for (int q = 0; q < 16; ++q) reinterpret_cast<uint4 *>(qrow)[q] = make_uint4(0, 0, 0, 0);
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_s, 128);
  }
}

}  // namespace attn
}  // namespace b200

using namespace b200;

extern "C" int32_t b200_launch_attention(const b200_tensor *q, const b200_tensor *k, const b200_tensor *v,
                                         const b200_tensor *mask, double scale, double mask_value, int32_t is_causal,
                                         const b200_tensor *out, const b200_tensor *weights, b200_stream s) {
  B200_REQUIRE(q && k && v && out, B200_ERR_INVALID, "null argument");
  for (const b200_tensor *t : {q, k, v, out})
    B200_REQUIRE(t->rank == 4 && t->dtype == B200_F32 && t->ptr, B200_ERR_UNSUPPORTED, "attention operands must be f32 [B, H, S, D]");
  const int64_t B = q->shape[0], H = q->shape[1], Sq = q->shape[2], D = q->shape[3], Sk = k->shape[2], Dv = v->shape[3];
  B200_REQUIRE(k->shape[0] == B && k->shape[1] == H && k->shape[3] == D && v->shape[0] == B && v->shape[1] == H &&
                   v->shape[2] == Sk && out->shape[0] == B && out->shape[1] == H && out->shape[2] == Sq && out->shape[3] == Dv,
               B200_ERR_SHAPE, "attention shape mismatch");
  B200_REQUIRE(D == attn::HD && Dv == attn::HD, B200_ERR_UNSUPPORTED,
               "the fused attention kernel is built for head dim 64 (got %lld / %lld); use the op chain", (long long)D, (long long)Dv);
  if (B == 0 || H == 0 || Sq == 0) return B200_OK;
  B200_REQUIRE(Sk > 0 && Sq < (1ll << 30) && Sk < (1ll << 30), B200_ERR_UNSUPPORTED, "attention sequence lengths out of range");
  attn::Params P;
  memset(&P, 0, sizeof(P));
  auto operand = [&](const b200_tensor *t, bool mn_major, Operand &o) -> int32_t {
    o.ptr = t->ptr;
    o.es = 4;
    o.mn_major = mn_major;
    o.s_mn = mn_major ? t->strides[3] : t->strides[2];
    o.s_k = mn_major ? t->strides[2] : t->strides[3];
    o.s_b[0] = 0; o.s_b[1] = t->strides[0]; o.s_b[2] = t->strides[1];
    o.bsz[0] = 1; o.bsz[1] = (int32_t)B; o.bsz[2] = (int32_t)H;
    B200_REQUIRE(t->strides[3] == 1, B200_ERR_UNSUPPORTED, "attention operands need a contiguous head dim");
    const int64_t mn = mn_major ? t->shape[3] : t->shape[2], kk = mn_major ? t->shape[2] : t->shape[3];
    B200_REQUIRE(tma_ok(o, mn, kk), B200_ERR_UNSUPPORTED, "attention operand strides must be multiples of 4 elements, 16-byte aligned");
    return B200_OK;
  };
  Operand oq, ok_, ov;
  int32_t st;
  if ((st = operand(q, false, oq)) != B200_OK) return st;
  if ((st = operand(k, false, ok_)) != B200_OK) return st;
  if ((st = operand(v, true, ov)) != B200_OK) return st;
  B200_REQUIRE(((uintptr_t)q->ptr | (uintptr_t)k->ptr | (uintptr_t)v->ptr | (uintptr_t)out->ptr) % 16 == 0, B200_ERR_UNSUPPORTED,
               "attention operands must be 16-byte aligned");
  if ((st = make_tmap(&P.tma_q, oq, Sq, D, attn::BQ)) != B200_OK) return st;
  if ((st = make_tmap(&P.tma_k, ok_, Sk, D, attn::BKV)) != B200_OK) return st;
  if ((st = make_tmap(&P.tma_v, ov, Dv, Sk)) != B200_OK) return st;
  B200_REQUIRE(out->strides[3] == 1 && out->strides[0] % 4 == 0 && out->strides[1] % 4 == 0 && out->strides[2] % 4 == 0,
               B200_ERR_UNSUPPORTED, "attention output needs a contiguous head dim and 16-byte-multiple strides");
  P.out = reinterpret_cast<float *>(out->ptr);
  P.o_sb = out->strides[0]; P.o_sh = out->strides[1]; P.o_ss = out->strides[2];
  if (weights) {
    B200_REQUIRE(weights->rank == 4 && weights->dtype == B200_F32 && weights->ptr && weights->shape[0] == B &&
                     weights->shape[1] == H && weights->shape[2] == Sq && weights->shape[3] == Sk,
                 B200_ERR_SHAPE, "attention weights must be f32 [B, H, Sq, Sk]");
    B200_REQUIRE(weights->strides[3] == 1 && Sk % 4 == 0 && ((uintptr_t)weights->ptr % 16) == 0 && weights->strides[0] % 4 == 0 &&
                     weights->strides[1] % 4 == 0 && weights->strides[2] % 4 == 0,
                 B200_ERR_UNSUPPORTED, "attention weights need Sk %% 4 == 0 and 16-byte-multiple strides");
    P.w = reinterpret_cast<float *>(weights->ptr);
    P.w_sb = weights->strides[0]; P.w_sh = weights->strides[1]; P.w_ss = weights->strides[2];
    EncodeTiledFn enc = encode_fn();
    B200_REQUIRE(enc, B200_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable from the driver");
    cuuint64_t dims[5] = {(cuuint64_t)Sk, (cuuint64_t)Sq, (cuuint64_t)H, (cuuint64_t)B, 1};
    cuuint64_t strides[4] = {(cuuint64_t)std::max<int64_t>(P.w_ss, 4) * 4, (cuuint64_t)std::max<int64_t>(P.w_sh, 4) * 4,
                             (cuuint64_t)std::max<int64_t>(P.w_sb, 4) * 4, 16};
    cuuint32_t box[5] = {32, (cuuint32_t)attn::BQ, 1, 1, 1}, estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&P.tma_w, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, weights->ptr, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_REQUIRE(r == CUDA_SUCCESS, B200_ERR_CUDA, "cuTensorMapEncodeTiled (attention weights) failed with %d", (int)r);
  }
  if (mask) {
    B200_REQUIRE(mask->rank == 4 && (mask->dtype == B200_BOOL || mask->dtype == B200_U8) && mask->ptr, B200_ERR_UNSUPPORTED,
                 "attention mask must be bool [B|1, H|1, Sq, Sk]");
    B200_REQUIRE((mask->shape[0] == B || mask->shape[0] == 1) && (mask->shape[1] == H || mask->shape[1] == 1) &&
                     mask->shape[2] == Sq && mask->shape[3] == Sk,
                 B200_ERR_SHAPE, "attention mask is not broadcastable to [B, H, Sq, Sk]");
    B200_REQUIRE(mask->strides[3] == 1 && Sk % 4 == 0 && ((uintptr_t)mask->ptr % 4) == 0 && mask->strides[2] % 4 == 0 &&
                     mask->strides[0] % 4 == 0 && mask->strides[1] % 4 == 0,
                 B200_ERR_UNSUPPORTED, "attention mask needs Sk %% 4 == 0 and 4-byte-multiple strides");
    P.mask = reinterpret_cast<const uint8_t *>(mask->ptr);
    P.m_sb = mask->shape[0] == 1 ? 0 : mask->strides[0];
    P.m_sh = mask->shape[1] == 1 ? 0 : mask->strides[1];
    P.m_ss = mask->strides[2];
    P.m_bcast_b = mask->shape[0] == 1;
    P.m_bcast_h = mask->shape[1] == 1;
    // 16-byte-multiple strides: mask tiles go through TMA (coalesced, zero-filled past the edges)
    if (((uintptr_t)mask->ptr % 16) == 0 && P.m_ss % 16 == 0 && P.m_sb % 16 == 0 && P.m_sh % 16 == 0) {
      EncodeTiledFn enc = encode_fn();
      B200_REQUIRE(enc, B200_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable from the driver");
      cuuint64_t dims[5] = {(cuuint64_t)Sk, (cuuint64_t)Sq, (cuuint64_t)(P.m_bcast_h ? 1 : H), (cuuint64_t)(P.m_bcast_b ? 1 : B), 1};
      cuuint64_t strides[4] = {(cuuint64_t)std::max<int64_t>(P.m_ss, 16), (cuuint64_t)(P.m_bcast_h ? 16 : P.m_sh),
                               (cuuint64_t)(P.m_bcast_b ? 16 : P.m_sb), 16};
      cuuint32_t box[5] = {(cuuint32_t)attn::BKV, (cuuint32_t)attn::BQ, 1, 1, 1}, estr[5] = {1, 1, 1, 1, 1};
      CUresult r = enc(&P.tma_m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 5, mask->ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                       CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r == CUDA_SUCCESS) P.mask_tma = 1;
    }
  }
  P.B = (int32_t)B; P.H = (int32_t)H; P.Sq = (int32_t)Sq; P.Sk = (int32_t)Sk;
  P.causal = is_causal ? 1 : 0;
  P.q_blocks = (int32_t)((Sq + attn::BQ - 1) / attn::BQ);
  P.scale = (float)scale;
  P.mask_value = (float)mask_value;
  const int64_t ctas = B * H * P.q_blocks;
  B200_REQUIRE(ctas < (1ll << 31), B200_ERR_UNSUPPORTED, "attention grid too large");
  const size_t smem = 1024 + attn::kQBytes + attn::kKBytes + attn::kVBytes + attn::kPBytes + (P.mask_tma ? attn::kMBytes : 0) + 128;
  auto launch = [&](auto kern) -> int32_t {
    // two CTAs per SM need the largest shared-memory carveout
    { const int32_t est = ensure_dyn_smem(reinterpret_cast<const void *>(kern), smem, true); if (est != B200_OK) return est; }
    kern<<<(unsigned)ctas, attn::kThreads, smem, resolve_stream(s)>>>(P);
    B200_LAUNCH_CHECK();
    return B200_OK;
  };
  st = P.mask_tma ? launch(attn::attention_fwd_kernel<true>) : launch(attn::attention_fwd_kernel<false>);
  if (st != B200_OK) return st;
  B200_LAUNCH_CHECK();
  return B200_OK;
}

extern "C" int32_t b200_launch_attention_backward(const b200_tensor *d_out, const b200_tensor *k, const b200_tensor *v,
                                                  const b200_tensor *out, const b200_tensor *weights, double scale,
                                                  int32_t is_causal, const b200_tensor *dq, const b200_tensor *ds,
                                                  b200_stream s) {
  B200_REQUIRE(d_out && k && v && out && weights && dq && ds, B200_ERR_INVALID, "null argument");
  for (const b200_tensor *t : {d_out, k, v, out, weights, dq, ds})
    B200_REQUIRE(t->rank == 4 && t->dtype == B200_F32 && t->ptr, B200_ERR_UNSUPPORTED, "attention operands must be f32 rank-4");
  const int64_t B = d_out->shape[0], H = d_out->shape[1], Sq = d_out->shape[2], D = d_out->shape[3], Sk = k->shape[2];
  B200_REQUIRE(D == attn::HD && k->shape[3] == attn::HD && v->shape[3] == attn::HD, B200_ERR_UNSUPPORTED,
               "the fused attention kernels are built for head dim 64; use the op chain");
  auto same = [&](const b200_tensor *t, int64_t s2, int64_t s3) {
    return t->shape[0] == B && t->shape[1] == H && t->shape[2] == s2 && t->shape[3] == s3;
  };
  B200_REQUIRE(same(k, Sk, D) && same(v, Sk, D) && same(out, Sq, D) && same(dq, Sq, D) && same(weights, Sq, Sk) && same(ds, Sq, Sk),
               B200_ERR_SHAPE, "attention backward shape mismatch");
  if (B == 0 || H == 0 || Sq == 0) return B200_OK;
  B200_REQUIRE(Sk > 0 && Sk % 4 == 0, B200_ERR_UNSUPPORTED, "attention backward needs Sk %% 4 == 0");
  attn::Params P;
  memset(&P, 0, sizeof(P));
  auto operand = [&](const b200_tensor *t, bool mn_major, Operand &o) -> int32_t {
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
  };
  Operand og, ov, ok_;
  int32_t st;
  if ((st = operand(d_out, false, og)) != B200_OK) return st;     // dO: A of dP
  if ((st = operand(v, false, ov)) != B200_OK) return st;         // V: K-major B of dP = dO·Vᵀ
  if ((st = operand(k, true, ok_)) != B200_OK) return st;         // K: MN-major B of dQ = dS·K
  if ((st = make_tmap(&P.tma_q, og, Sq, D, attn::BQ)) != B200_OK) return st;
  if ((st = make_tmap(&P.tma_k, ov, Sk, D, attn::BKV)) != B200_OK) return st;
  if ((st = make_tmap(&P.tma_v, ok_, D, Sk)) != B200_OK) return st;
  EncodeTiledFn enc = encode_fn();
  B200_REQUIRE(enc, B200_ERR_CUDA, "cuTensorMapEncodeTiled is unavailable from the driver");
  auto score_map = [&](const b200_tensor *t, CUtensorMap *map) -> int32_t {
    B200_REQUIRE(t->strides[3] == 1 && ((uintptr_t)t->ptr % 16) == 0 && t->strides[0] % 4 == 0 && t->strides[1] % 4 == 0 &&
                     t->strides[2] % 4 == 0,
                 B200_ERR_UNSUPPORTED, "attention weights / dS need 16-byte-multiple strides");
    cuuint64_t dims[5] = {(cuuint64_t)Sk, (cuuint64_t)Sq, (cuuint64_t)H, (cuuint64_t)B, 1};
    cuuint64_t strides[4] = {(cuuint64_t)std::max<int64_t>(t->strides[2], 4) * 4, (cuuint64_t)std::max<int64_t>(t->strides[1], 4) * 4,
                             (cuuint64_t)std::max<int64_t>(t->strides[0], 4) * 4, 16};
    cuuint32_t box[5] = {32, (cuuint32_t)attn::BQ, 1, 1, 1}, estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, t->ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_REQUIRE(r == CUDA_SUCCESS, B200_ERR_CUDA, "cuTensorMapEncodeTiled (attention scores) failed with %d", (int)r);
    return B200_OK;
  };
  if ((st = score_map(weights, &P.tma_p)) != B200_OK) return st;
  if ((st = score_map(ds, &P.tma_w)) != B200_OK) return st;
  auto rows_ok = [](const b200_tensor *t) {
    return t->strides[3] == 1 && ((uintptr_t)t->ptr % 16) == 0 && t->strides[0] % 4 == 0 && t->strides[1] % 4 == 0 && t->strides[2] % 4 == 0;
  };
  B200_REQUIRE(rows_ok(out) && rows_ok(dq) && rows_ok(d_out), B200_ERR_UNSUPPORTED, "attention rows need 16-byte-multiple strides");
  P.out = reinterpret_cast<float *>(dq->ptr);
  P.o_sb = dq->strides[0]; P.o_sh = dq->strides[1]; P.o_ss = dq->strides[2];
  P.w = reinterpret_cast<float *>(ds->ptr);
  P.w_sb = ds->strides[0]; P.w_sh = ds->strides[1]; P.w_ss = ds->strides[2];
  P.o_fwd = reinterpret_cast<const float *>(out->ptr);
  P.of_sb = out->strides[0]; P.of_sh = out->strides[1]; P.of_ss = out->strides[2];
  P.d_out = reinterpret_cast<const float *>(d_out->ptr);
  P.do_sb = d_out->strides[0]; P.do_sh = d_out->strides[1]; P.do_ss = d_out->strides[2];
  P.B = (int32_t)B; P.H = (int32_t)H; P.Sq = (int32_t)Sq; P.Sk = (int32_t)Sk;
  P.causal = is_causal ? 1 : 0;
  P.q_blocks = (int32_t)((Sq + attn::BQ - 1) / attn::BQ);
  P.scale = (float)scale;
  const int64_t ctas = B * H * P.q_blocks;
  B200_REQUIRE(ctas < (1ll << 31), B200_ERR_UNSUPPORTED, "attention grid too large");
  const size_t smem = 1024 + attn::kQBytes + attn::kKBytes + attn::kVBytes + attn::kPBytes + 128;
  { const int32_t est = ensure_dyn_smem(reinterpret_cast<const void *>(attn::attention_bwd_dq_kernel), smem, true); if (est != B200_OK) return est; }
  attn::attention_bwd_dq_kernel<<<(unsigned)ctas, attn::kThreads, smem, resolve_stream(s)>>>(P);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
