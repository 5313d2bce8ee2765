// Register-resident op-tape interpreter shared by the three kernel families
// (fused elementwise, fuse-on-read/-write reductions, GEMM epilogue).
//
// Behavioural spec: the FuseOp list a TraceOperationFuser builds
// (crates/burn-cubecl-fusion/src/engine/codegen/ir.rs:135-195) executed by
// fuse_on_write/fuse_on_read (crates/burn-cubecl-fusion/src/engine/codegen/kernel.rs:29-50).
// Numerics follow the CPU oracle burn-ndarray (crates/burn-ndarray/src/ops/tensor.rs):
// f32 IEEE add/sub/mul/div/sqrt with no contraction across ops; erf, tanh and
// the trigonometric family evaluated in f64 then rounded to f32 (:515-523,
// :620-626,:714-720); exp/log/log1p/powf in f32.
//
// Design (B200).  The public tape (b200_tape_op: three free operands) is compiled
// on the host (tape_host.cuh) into an accumulator ISA:
//        acc = OP(acc, B [, C])        B, C always live in shared memory
// Every thread owns U vectors of VEC consecutive elements; `acc` stays in
// registers for the whole tape.  Inputs and saved temporaries live in a
// thread-private slot file laid out [slot][u][thread] as 16-byte words, so every
// LDS.128/STS.128 is conflict-free and no barrier is needed; scalars are one
// CTA-shared broadcast word each.  Operand fetch is therefore branch-free
// (one address computation + U LDS.128) and opcode dispatch is warp-uniform
// (the program sits in the kernel-parameter constant bank): ~15 issue slots of
// overhead per op for 4*U elements of work per thread.
#pragma once
#include "common.cuh"
#include "tape_math.cuh"

namespace b200 {

constexpr int kMaxDims = B200_MAX_RANK;
constexpr int kMaxIOps = 160;  // internal ops after compilation (public limit 64)

enum LoadMode : int32_t {
  kModeVec = 0,     // VEC consecutive elements, aligned vector access
  kModeBcast = 1,   // innermost stride 0: one element broadcast to all lanes
  kModeGather = 2   // VEC element accesses at offset + j*inner_stride
};

// 64-bit internal op: {op:8, dst_tmp:8, dst_out:8, flags:8, b_addr:16, c_addr:16}
// b_addr/c_addr are in 16-byte (VEC==4) or 4-byte (VEC==1) words.
constexpr uint32_t kFlagBInput = 1u, kFlagBShared = 2u, kFlagCInput = 4u, kFlagCShared = 8u, kFlagHasB = 16u;

struct OperandDesc {
  void *ptr;
  int32_t s3[3];     // element strides of the (right-aligned) 3-D collapsed layout, rank <= 3
  int32_t dtype;
  int32_t mode;
  int32_t async_es;  // element size (1/2/4) when the operand streams through cp.async, else 0
  int64_t strides[kMaxDims];  // generic path (rank > 3): element strides, collapsed dims
};

// How a kernel instance turns a vector index into operand offsets.
enum RankMode : int {
  kRankLinear = 0,  // collapsed rank 1: offset = v * VEC * s3[2]
  kRank3 = 1,       // collapsed rank <= 3: two fast divisions, 32-bit math
  kRankGeneric = 2  // any rank: division loop, 64-bit math
};

struct TapeParams {
  uint64_t iops[kMaxIOps];
  uint32_t scalars[B200_MAX_TAPE_SCALARS];
  OperandDesc in[B200_MAX_TAPE_INPUTS];
  OperandDesc out[B200_MAX_TAPE_OUTPUTS];
  // Collapsed geometry.  rank <= 3 is stored right-aligned in 3 dims (leading
  // dims of size 1) so coordinate math is branch-free; larger ranks use the
  // generic loop.  div[k] divides by shape[k] (innermost measured in vectors).
  FastDiv div[kMaxDims];
  uint32_t shape[kMaxDims];
  int32_t n_iops, n_in, n_out, n_tmp, n_scalars, rank;
  uint32_t n_vec;  // number of VEC-wide vectors (numel / VEC)
};

// ----------------------------------------------------------------- typed vector IO
// Loads VEC consecutive elements of `dtype` starting at element offset `off`
// as 32-bit lanes (f32 bits for float types, i32 for int/bool types).
__device__ __forceinline__ void load_vec4(const void *base, int32_t dtype, int64_t off,
                                          uint32_t (&r)[4]) {
  switch (dtype) {
    case B200_F32:
    case B200_I32: {
      const uint4 v = __ldcs(reinterpret_cast<const uint4 *>(
          reinterpret_cast<const uint32_t *>(base) + off));
      r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
      break;
    }
    case B200_BF16: {
      const uint2 v = __ldcs(reinterpret_cast<const uint2 *>(
          reinterpret_cast<const uint16_t *>(base) + off));
      r[0] = v.x << 16; r[1] = v.x & 0xFFFF0000u;
      r[2] = v.y << 16; r[3] = v.y & 0xFFFF0000u;
      break;
    }
    case B200_F16: {
      const uint2 v = __ldcs(reinterpret_cast<const uint2 *>(
          reinterpret_cast<const uint16_t *>(base) + off));
      const __half2 h0 = *reinterpret_cast<const __half2 *>(&v.x);
      const __half2 h1 = *reinterpret_cast<const __half2 *>(&v.y);
      const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
      r[0] = u_of(f0.x); r[1] = u_of(f0.y); r[2] = u_of(f1.x); r[3] = u_of(f1.y);
      break;
    }
    case B200_I64: {
      const longlong2 *p = reinterpret_cast<const longlong2 *>(
          reinterpret_cast<const int64_t *>(base) + off);
      const longlong2 v0 = __ldcs(p), v1 = __ldcs(p + 1);
      r[0] = (uint32_t)v0.x; r[1] = (uint32_t)v0.y;
      r[2] = (uint32_t)v1.x; r[3] = (uint32_t)v1.y;
      break;
    }
    default: {  // BOOL / U8
      const uint32_t v = __ldcs(reinterpret_cast<const uint32_t *>(
          reinterpret_cast<const uint8_t *>(base) + off));
      r[0] = v & 0xFFu; r[1] = (v >> 8) & 0xFFu; r[2] = (v >> 16) & 0xFFu; r[3] = v >> 24;
      break;
    }
  }
}

__device__ __forceinline__ uint32_t load_one(const void *base, int32_t dtype, int64_t off) {
  switch (dtype) {
    case B200_F32:
    case B200_I32: return __ldg(reinterpret_cast<const uint32_t *>(base) + off);
    case B200_BF16:
      return (uint32_t)__ldg(reinterpret_cast<const uint16_t *>(base) + off) << 16;
    case B200_F16:
      return u_of(__half2float(__ushort_as_half(__ldg(reinterpret_cast<const uint16_t *>(base) + off))));
    case B200_I64: return (uint32_t)__ldg(reinterpret_cast<const long long *>(base) + off);
    default: return (uint32_t)__ldg(reinterpret_cast<const uint8_t *>(base) + off);
  }
}

__device__ __forceinline__ void store_one(void *base, int32_t dtype, int64_t off, uint32_t v) {
  switch (dtype) {
    case B200_F32:
    case B200_I32: reinterpret_cast<uint32_t *>(base)[off] = v; break;
    case B200_BF16:
      reinterpret_cast<__nv_bfloat16 *>(base)[off] = __float2bfloat16_rn(f_of(v));
      break;
    case B200_F16: reinterpret_cast<__half *>(base)[off] = __float2half_rn(f_of(v)); break;
    case B200_I64: reinterpret_cast<long long *>(base)[off] = (long long)(int32_t)v; break;
    default: reinterpret_cast<uint8_t *>(base)[off] = (uint8_t)v; break;
  }
}

__device__ __forceinline__ void store_vec4(void *base, int32_t dtype, int64_t off,
                                           const uint32_t (&r)[4]) {
  switch (dtype) {
    case B200_F32:
    case B200_I32:
      __stcs(reinterpret_cast<uint4 *>(reinterpret_cast<uint32_t *>(base) + off),
             make_uint4(r[0], r[1], r[2], r[3]));
      break;
    case B200_BF16: {
      __nv_bfloat162 lo = __floats2bfloat162_rn(f_of(r[0]), f_of(r[1]));
      __nv_bfloat162 hi = __floats2bfloat162_rn(f_of(r[2]), f_of(r[3]));
      uint2 v;
      v.x = *reinterpret_cast<uint32_t *>(&lo);
      v.y = *reinterpret_cast<uint32_t *>(&hi);
      __stcs(reinterpret_cast<uint2 *>(reinterpret_cast<uint16_t *>(base) + off), v);
      break;
    }
    case B200_F16: {
      __half2 lo = __floats2half2_rn(f_of(r[0]), f_of(r[1]));
      __half2 hi = __floats2half2_rn(f_of(r[2]), f_of(r[3]));
      uint2 v;
      v.x = *reinterpret_cast<uint32_t *>(&lo);
      v.y = *reinterpret_cast<uint32_t *>(&hi);
      __stcs(reinterpret_cast<uint2 *>(reinterpret_cast<uint16_t *>(base) + off), v);
      break;
    }
    case B200_I64: {
      longlong2 *p = reinterpret_cast<longlong2 *>(reinterpret_cast<int64_t *>(base) + off);
      __stcs(p, make_longlong2((long long)(int32_t)r[0], (long long)(int32_t)r[1]));
      __stcs(p + 1, make_longlong2((long long)(int32_t)r[2], (long long)(int32_t)r[3]));
      break;
    }
    default: {
      const uint32_t v = (r[0] & 0xFFu) | ((r[1] & 0xFFu) << 8) | ((r[2] & 0xFFu) << 16) |
                         (r[3] << 24);
      __stcs(reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(base) + off), v);
      break;
    }
  }
}

// ----------------------------------------------------------------- coordinates
// Collapsed coordinates of vector index v for the rank <= 3 layout (right-aligned;
// the innermost coordinate is in elements).
struct Coord3 {
  uint32_t c0, c1, c2;
};

template <int VEC, int RM>
__device__ __forceinline__ Coord3 coords3(const TapeParams &p, uint32_t v) {
  Coord3 c;
  if constexpr (RM == kRank3) {
    const uint32_t t = fd_div(v, p.div[2]);
    c.c2 = (v - t * p.div[2].d) * VEC;
    c.c0 = fd_div(t, p.div[1]);
    c.c1 = t - c.c0 * p.div[1].d;
  } else {
    c.c0 = c.c1 = 0;
    c.c2 = v * VEC;
  }
  return c;
}

// Element offset of operand `d` for vector v.
template <int VEC, int RM>
__device__ __forceinline__ int64_t operand_offset(const TapeParams &p, const OperandDesc &d,
                                                  uint32_t v, const Coord3 &c) {
  if constexpr (RM == kRankLinear) {
    return (int64_t)(int32_t)(c.c2 * (uint32_t)d.s3[2]);
  } else if constexpr (RM == kRank3) {
    return (int64_t)(int32_t)(c.c0 * (uint32_t)d.s3[0] + c.c1 * (uint32_t)d.s3[1] + c.c2 * (uint32_t)d.s3[2]);
  } else {
    int64_t off = 0;
    uint32_t rest = v;
    for (int k = p.rank - 1; k > 0; --k) {
      const uint32_t q = fd_div(rest, p.div[k]);
      const uint32_t r = rest - q * p.div[k].d;
      off += (int64_t)(k == p.rank - 1 ? r * VEC : r) * d.strides[k];
      rest = q;
    }
    return off + (int64_t)(p.rank == 1 ? rest * VEC : rest) * d.strides[0];
  }
}

template <int RM>
__device__ __forceinline__ int64_t inner_stride(const TapeParams &p, const OperandDesc &d) {
  if constexpr (RM == kRankGeneric) return d.strides[p.rank - 1];
  else return (int64_t)d.s3[2];
}

// Synchronous load of the VEC lanes of one operand.
template <int VEC, int RM>
__device__ __forceinline__ void load_operand(const TapeParams &p, const OperandDesc &d, uint32_t v,
                                             const Coord3 &c, uint32_t (&r)[VEC]) {
  const int64_t off = operand_offset<VEC, RM>(p, d, v, c);
  if constexpr (VEC == 4) {
    if (d.mode == kModeVec) {
      load_vec4(d.ptr, d.dtype, off, r);
      return;
    }
  }
  if (d.mode == kModeBcast) {
    const uint32_t x = load_one(d.ptr, d.dtype, off);
#pragma unroll
    for (int j = 0; j < VEC; ++j) r[j] = x;
  } else {
    const int64_t s = inner_stride<RM>(p, d);
#pragma unroll
    for (int j = 0; j < VEC; ++j) r[j] = load_one(d.ptr, d.dtype, off + j * s);
  }
}

template <int VEC, int RM>
__device__ __forceinline__ void store_operand(const TapeParams &p, const OperandDesc &d, uint32_t v,
                                              const Coord3 &c, const uint32_t (&r)[VEC]) {
  const int64_t off = operand_offset<VEC, RM>(p, d, v, c);
  if constexpr (VEC == 4) {
    if (d.mode == kModeVec) {
      store_vec4(d.ptr, d.dtype, off, r);
      return;
    }
  }
  const int64_t s = inner_stride<RM>(p, d);
#pragma unroll
  for (int j = 0; j < VEC; ++j) store_one(d.ptr, d.dtype, off + j * s, r[j]);
}

// ----------------------------------------------------------------- async copies
__device__ __forceinline__ void cp_async_16(uint32_t smem_addr, const void *g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_addr), "l"(g));
}
__device__ __forceinline__ void cp_async_8(uint32_t smem_addr, const void *g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_addr), "l"(g));
}
__device__ __forceinline__ void cp_async_4(uint32_t smem_addr, const void *g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_addr), "l"(g));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
// Waits until at most `n` commit groups are pending (n is warp-uniform, 0..3).
__device__ __forceinline__ void cp_async_wait_dyn(int n) {
  switch (n) {
    case 0: asm volatile("cp.async.wait_group 0;\n" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;\n" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;\n" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 3;\n" ::: "memory"); break;
  }
}

// Expands a raw 16-byte slot word filled by cp.async (8/16-bit element types)
// into 32-bit lanes.
__device__ __forceinline__ void expand_raw(int32_t dtype, uint32_t (&r)[4]) {
  if (dtype == B200_BF16) {
    const uint32_t x = r[0], y = r[1];
    r[0] = x << 16; r[1] = x & 0xFFFF0000u; r[2] = y << 16; r[3] = y & 0xFFFF0000u;
  } else if (dtype == B200_F16) {
    const uint32_t x = r[0], y = r[1];
    const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&x));
    const float2 f1 = __half22float2(*reinterpret_cast<const __half2 *>(&y));
    r[0] = u_of(f0.x); r[1] = u_of(f0.y); r[2] = u_of(f1.x); r[3] = u_of(f1.y);
  } else {  // BOOL / U8
    const uint32_t v = r[0];
    r[0] = v & 0xFFu; r[1] = (v >> 8) & 0xFFu; r[2] = (v >> 16) & 0xFFu; r[3] = v >> 24;
  }
}

// ----------------------------------------------------------------- slot file
// Word (slot, u) of thread t sits at word index (slot*U + u)*BLOCK + t, where a
// word is 16 bytes (VEC == 4) or 4 bytes (VEC == 1).  Shared scalar words follow
// the thread-private region.
template <int VEC, int U, int BLOCK>
struct SlotFile {
  uint32_t *smem;  // base of the slot file (word 0)
  int tid;
  static constexpr int kWordU32 = VEC;  // u32 per word

  __device__ __forceinline__ uint32_t *word(uint32_t widx) const { return smem + (size_t)widx * kWordU32; }
  __device__ __forceinline__ uint32_t private_word(int slot, int u) const {
    return (uint32_t)((slot * U + u) * BLOCK + tid);
  }
  __device__ __forceinline__ void put_w(uint32_t widx, const uint32_t (&r)[VEC]) const {
    if constexpr (VEC == 4) *reinterpret_cast<uint4 *>(word(widx)) = make_uint4(r[0], r[1], r[2], r[3]);
    else *word(widx) = r[0];
  }
  __device__ __forceinline__ void get_w(uint32_t widx, uint32_t (&r)[VEC]) const {
    if constexpr (VEC == 4) {
      const uint4 v = *reinterpret_cast<const uint4 *>(word(widx));
      r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
    } else {
      r[0] = *word(widx);
    }
  }
  __device__ __forceinline__ void put(int slot, int u, const uint32_t (&r)[VEC]) const {
    put_w(private_word(slot, u), r);
  }
  __device__ __forceinline__ void get(int slot, int u, uint32_t (&r)[VEC]) const {
    get_w(private_word(slot, u), r);
  }
};

// Bytes of a slot file with `n_slots` private slots and `n_scalars` shared words.
static inline size_t slot_file_bytes(int n_slots, int n_scalars, int vec, int u, int block) {
  return ((size_t)n_slots * u * block + (size_t)n_scalars) * vec * 4;
}

// Writes the CTA-shared scalar words.  Callers __syncthreads() afterwards.
template <int VEC, int U, int BLOCK>
__device__ __forceinline__ void init_scalars(const TapeParams &p, const SlotFile<VEC, U, BLOCK> &slots,
                                             int n_private_slots) {
  const uint32_t base = (uint32_t)(n_private_slots * U * BLOCK);
  for (int i = slots.tid; i < p.n_scalars; i += BLOCK) {
    uint32_t r[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) r[j] = p.scalars[i];
    slots.put_w(base + i, r);
  }
}

// ----------------------------------------------------------------- interpreter
#define B200_UN(expr)                                   \
  _Pragma("unroll") for (int u = 0; u < U; ++u)         \
  _Pragma("unroll") for (int j = 0; j < VEC; ++j) {     \
    const float x = f_of(acc[u][j]);                    \
    (void)x;                                            \
    acc[u][j] = u_of(expr);                             \
  }
#define B200_BIN(expr)                                  \
  _Pragma("unroll") for (int u = 0; u < U; ++u)         \
  _Pragma("unroll") for (int j = 0; j < VEC; ++j) {     \
    const float x = f_of(acc[u][j]), y = f_of(b[u][j]); \
    acc[u][j] = u_of(expr);                             \
  }
#define B200_CMP(expr)                                  \
  _Pragma("unroll") for (int u = 0; u < U; ++u)         \
  _Pragma("unroll") for (int j = 0; j < VEC; ++j) {     \
    const float x = f_of(acc[u][j]), y = f_of(b[u][j]); \
    (void)y;                                            \
    acc[u][j] = (expr) ? 1u : 0u;                       \
  }
#define B200_IUN(expr)                                  \
  _Pragma("unroll") for (int u = 0; u < U; ++u)         \
  _Pragma("unroll") for (int j = 0; j < VEC; ++j) {     \
    const int32_t x = (int32_t)acc[u][j];               \
    (void)x;                                            \
    acc[u][j] = (uint32_t)(expr);                       \
  }
#define B200_IBIN(expr)                                           \
  _Pragma("unroll") for (int u = 0; u < U; ++u)                   \
  _Pragma("unroll") for (int j = 0; j < VEC; ++j) {               \
    const int32_t x = (int32_t)acc[u][j], y = (int32_t)b[u][j];   \
    (void)y;                                                      \
    acc[u][j] = (uint32_t)(expr);                                 \
  }

// Number of operands each public opcode reads (1, 2 or 3).
static inline int op_arity(int op) {
  if (op == B200_OP_CLAMP_F || op == B200_OP_CLAMP_I || op == B200_OP_SELECT) return 3;
  switch (op) {
    case B200_OP_ADD_F: case B200_OP_SUB_F: case B200_OP_MUL_F: case B200_OP_DIV_F:
    case B200_OP_REM_F: case B200_OP_POW_F: case B200_OP_MIN_F: case B200_OP_MAX_F:
    case B200_OP_ATAN2_F: case B200_OP_REMT_F:
    case B200_OP_EQ_F: case B200_OP_NE_F: case B200_OP_LT_F: case B200_OP_LE_F:
    case B200_OP_GT_F: case B200_OP_GE_F:
    case B200_OP_ADD_I: case B200_OP_SUB_I: case B200_OP_MUL_I: case B200_OP_DIV_I:
    case B200_OP_REM_I: case B200_OP_MIN_I: case B200_OP_MAX_I: case B200_OP_AND_I:
    case B200_OP_OR_I: case B200_OP_XOR_I: case B200_OP_SHL_I: case B200_OP_SHR_I:
    case B200_OP_EQ_I: case B200_OP_NE_I: case B200_OP_LT_I: case B200_OP_LE_I:
    case B200_OP_GT_I: case B200_OP_GE_I:
    case B200_OP_AND_B: case B200_OP_OR_B: case B200_OP_XOR_B:
      return 2;
    default: return 1;
  }
}

// Fetches operand words for all U vectors.  `waddr` is the operand's word
// address for u == 0 relative to the slot-file base; private operands add the
// thread id and stride BLOCK words per u, shared ones are broadcast.
template <int VEC, int U, int BLOCK>
__device__ __forceinline__ void fetch_operand(const SlotFile<VEC, U, BLOCK> &slots, uint32_t waddr,
                                              bool is_input, bool is_shared, uint32_t stage_off,
                                              uint32_t (&dst)[U][VEC]) {
  const uint32_t base = waddr + (is_input ? stage_off : 0u) + (is_shared ? 0u : (uint32_t)slots.tid);
  const uint32_t ustride = is_shared ? 0u : (uint32_t)BLOCK;
#pragma unroll
  for (int u = 0; u < U; ++u) slots.get_w(base + u * ustride, dst[u]);
}

// Runs the compiled tape.  Input slots of the current ring stage start at
// private slot `in_base`; temporaries at `tmp_base`.  `store(out_index, acc)` is
// called for every op carrying a dst_out.
template <int VEC, int U, int BLOCK, typename StoreFn>
__device__ __forceinline__ void run_tape(const TapeParams &p, const SlotFile<VEC, U, BLOCK> &slots,
                                         uint32_t (&acc)[U][VEC], StoreFn &&store, int in_base,
                                         int tmp_base) {
  const uint32_t stage_off = (uint32_t)(in_base * U * BLOCK);
  for (int pc = 0; pc < p.n_iops; ++pc) {
    const uint64_t w = p.iops[pc];
    const uint32_t lo = (uint32_t)w, hi = (uint32_t)(w >> 32);
    const uint32_t opc = lo & 0xFFu, dst_tmp = (lo >> 8) & 0xFFu, dst_out = (lo >> 16) & 0xFFu;
    const uint32_t flags = lo >> 24;
    uint32_t b[U][VEC], c[U][VEC];
    if ((flags & kFlagHasB) && opc != kOpLoad)
      fetch_operand<VEC, U, BLOCK>(slots, hi & 0xFFFFu, flags & kFlagBInput, flags & kFlagBShared, stage_off, b);
    switch (opc) {
      case kOpLoad:  // acc = B, fetched straight into the accumulator registers
        fetch_operand<VEC, U, BLOCK>(slots, hi & 0xFFFFu, flags & kFlagBInput, flags & kFlagBShared, stage_off, acc);
        break;
      case kOpSave: break;
      case B200_OP_ADD_F: B200_BIN(__fadd_rn(x, y)) break;
      case B200_OP_SUB_F: B200_BIN(__fsub_rn(x, y)) break;
      case B200_OP_MUL_F: B200_BIN(__fmul_rn(x, y)) break;
      case B200_OP_DIV_F: B200_BIN(__fdiv_rn(x, y)) break;
      case kOpDivScalar: {
        const float ys = f_of(b[0][0]);
        const float rinv = __frcp_rn(ys);
        bool safe = true;
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int j = 0; j < VEC; ++j) safe &= div_scalar_safe(f_of(acc[u][j]));
        if (safe) {
          B200_UN(div_scalar_fast(x, ys, rinv))
        } else {
          B200_UN(__fdiv_rn(x, ys))
        }
        break;
      }
      case kOpSquare: B200_UN(__fmul_rn(x, x)) break;
      case kOpCube: B200_UN(__fmul_rn(__fmul_rn(x, x), x)) break;
      case kOpMulAdd:
        fetch_operand<VEC, U, BLOCK>(slots, hi >> 16, flags & kFlagCInput, flags & kFlagCShared, stage_off, c);
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int j = 0; j < VEC; ++j)
            acc[u][j] = u_of(__fadd_rn(__fmul_rn(f_of(acc[u][j]), f_of(b[u][j])), f_of(c[u][j])));
        break;
      case kOpGelu: {
        // div_scalar(sqrt2) -> erf -> add_scalar(1) -> mul(x, .) -> div_scalar(2), each
        // step rounded exactly as the five separate ops would be
        // (crates/burn-backend/src/backend/ops/activation.rs:69-76)
        const float s2 = 1.41421353816986083984375f;  // f32(SQRT_2)
        const float rinv = 0.707106769084930419921875f;  // RN(1 / f32(SQRT_2))
        if (!(flags & kFlagHasB)) {  // x is the accumulator itself
#pragma unroll
          for (int u = 0; u < U; ++u)
#pragma unroll
            for (int j = 0; j < VEC; ++j) b[u][j] = acc[u][j];
        }
        bool safe = true;
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int j = 0; j < VEC; ++j) safe &= div_scalar_safe(f_of(b[u][j]));
        if (safe) {
#pragma unroll
          for (int u = 0; u < U; ++u)
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
              const float x = f_of(b[u][j]);
              const float e = erf_f32(div_scalar_fast(x, s2, rinv));
              acc[u][j] = u_of(__fmul_rn(__fmul_rn(x, __fadd_rn(e, 1.0f)), 0.5f));
            }
        } else {
#pragma unroll
          for (int u = 0; u < U; ++u)
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
              const float x = f_of(b[u][j]);
              const float e = erf_f32(__fdiv_rn(x, s2));
              acc[u][j] = u_of(__fmul_rn(__fmul_rn(x, __fadd_rn(e, 1.0f)), 0.5f));
            }
        }
        break;
      }
      case B200_OP_REM_F: B200_BIN(rem_floor(x, y)) break;
      case B200_OP_REMT_F: B200_BIN(rem_tensor(x, y)) break;
      case B200_OP_POW_F: B200_BIN(pow_f(x, y)) break;
      case B200_OP_MIN_F: B200_BIN((x != x || y != y) ? __int_as_float(0x7fc00000) : fminf(x, y)) break;
      case B200_OP_MAX_F: B200_BIN((x != x || y != y) ? __int_as_float(0x7fc00000) : fmaxf(x, y)) break;
      case B200_OP_ATAN2_F: B200_BIN((float)atan2((double)x, (double)y)) break;
      case B200_OP_NEG_F: B200_UN(-x) break;
      case B200_OP_ABS_F: B200_UN(fabsf(x)) break;
      case B200_OP_EXP_F: B200_UN(expf(x)) break;
      case B200_OP_LOG_F: B200_UN(logf(x)) break;
      case B200_OP_LOG1P_F: B200_UN(log1pf(x)) break;
      case B200_OP_SQRT_F: B200_UN(__fsqrt_rn(x)) break;
      case B200_OP_RECIP_F: B200_UN(__fdiv_rn(1.0f, x)) break;
      case B200_OP_TANH_F: B200_UN(tanh_f32(x)) break;
      case B200_OP_ERF_F: B200_UN(erf_f32(x)) break;
      case B200_OP_FLOOR_F: B200_UN(floorf(x)) break;
      case B200_OP_CEIL_F: B200_UN(ceilf(x)) break;
      case B200_OP_ROUND_F: B200_UN(rintf(x)) break;
      case B200_OP_TRUNC_F: B200_UN(truncf(x)) break;
      case B200_OP_SIGN_F: B200_UN(sign_f(x)) break;
      case B200_OP_SIGMOID_F: B200_UN(__fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)))) break;
      case B200_OP_SIN_F: case B200_OP_COS_F: case B200_OP_TAN_F: case B200_OP_SINH_F:
      case B200_OP_COSH_F: case B200_OP_ASIN_F: case B200_OP_ACOS_F: case B200_OP_ATAN_F:
      case B200_OP_ASINH_F: case B200_OP_ACOSH_F: case B200_OP_ATANH_F:
        B200_UN(slow_unary((int)opc, x)) break;
      case B200_OP_CLAMP_F:
        fetch_operand<VEC, U, BLOCK>(slots, hi >> 16, flags & kFlagCInput, flags & kFlagCShared, stage_off, c);
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int j = 0; j < VEC; ++j) {
            const float x = f_of(acc[u][j]), lo_ = f_of(b[u][j]), hi_ = f_of(c[u][j]);
            acc[u][j] = u_of(x != x ? x : fminf(fmaxf(x, lo_), hi_));  // NaN stays NaN
          }
        break;
      case B200_OP_EQ_F: B200_CMP(x == y) break;
      case B200_OP_NE_F: B200_CMP(x != y) break;
      case B200_OP_LT_F: B200_CMP(x < y) break;
      case B200_OP_LE_F: B200_CMP(x <= y) break;
      case B200_OP_GT_F: B200_CMP(x > y) break;
      case B200_OP_GE_F: B200_CMP(x >= y) break;
      case B200_OP_ISNAN_F: B200_CMP(x != x) break;
      case B200_OP_ISINF_F: B200_CMP(isinf(x)) break;
      case B200_OP_ADD_I: B200_IBIN(x + y) break;
      case B200_OP_SUB_I: B200_IBIN(x - y) break;
      case B200_OP_MUL_I: B200_IBIN(x * y) break;
      case B200_OP_DIV_I: B200_IBIN(y == 0 ? 0 : x / y) break;
      case B200_OP_REM_I: B200_IBIN(irem_floor(x, y)) break;
      case B200_OP_MIN_I: B200_IBIN(min(x, y)) break;
      case B200_OP_MAX_I: B200_IBIN(max(x, y)) break;
      case B200_OP_NEG_I: B200_IUN(-x) break;
      case B200_OP_ABS_I: B200_IUN(abs(x)) break;
      case B200_OP_SIGN_I: B200_IUN((x > 0) - (x < 0)) break;
      case B200_OP_AND_I: B200_IBIN(x & y) break;
      case B200_OP_OR_I: B200_IBIN(x | y) break;
      case B200_OP_XOR_I: B200_IBIN(x ^ y) break;
      case B200_OP_NOT_I: B200_IUN(~x) break;
      case B200_OP_SHL_I: B200_IBIN(x << (y & 31)) break;
      case B200_OP_SHR_I: B200_IBIN(x >> (y & 31)) break;
      case B200_OP_CLAMP_I:
        fetch_operand<VEC, U, BLOCK>(slots, hi >> 16, flags & kFlagCInput, flags & kFlagCShared, stage_off, c);
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int j = 0; j < VEC; ++j)
            acc[u][j] = (uint32_t)min(max((int32_t)acc[u][j], (int32_t)b[u][j]), (int32_t)c[u][j]);
        break;
      case B200_OP_EQ_I: B200_IBIN(x == y ? 1 : 0) break;
      case B200_OP_NE_I: B200_IBIN(x != y ? 1 : 0) break;
      case B200_OP_LT_I: B200_IBIN(x < y ? 1 : 0) break;
      case B200_OP_LE_I: B200_IBIN(x <= y ? 1 : 0) break;
      case B200_OP_GT_I: B200_IBIN(x > y ? 1 : 0) break;
      case B200_OP_GE_I: B200_IBIN(x >= y ? 1 : 0) break;
      case B200_OP_AND_B: B200_IBIN((x != 0) & (y != 0)) break;
      case B200_OP_OR_B: B200_IBIN((x != 0) | (y != 0)) break;
      case B200_OP_XOR_B: B200_IBIN((x != 0) ^ (y != 0)) break;
      case B200_OP_NOT_B: B200_IUN(x == 0 ? 1 : 0) break;
      case B200_OP_SELECT:  // acc = C ? B : acc
        fetch_operand<VEC, U, BLOCK>(slots, hi >> 16, flags & kFlagCInput, flags & kFlagCShared, stage_off, c);
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int j = 0; j < VEC; ++j) acc[u][j] = c[u][j] ? b[u][j] : acc[u][j];
        break;
      case B200_OP_F2I: B200_UN(__int_as_float(__float2int_rz(x))) break;
      case B200_OP_I2F: B200_IUN(__float_as_int(__int2float_rn(x))) break;
      case B200_OP_B2F: B200_IUN(__float_as_int(x ? 1.0f : 0.0f)) break;
      case B200_OP_B2I: B200_IUN(x ? 1 : 0) break;
      case B200_OP_F2B: B200_UN(__int_as_float(x != 0.0f ? 1 : 0)) break;
      case B200_OP_I2B: B200_IUN(x != 0 ? 1 : 0) break;
      default: break;
    }
    if (dst_tmp != B200_DST_NONE) {
#pragma unroll
      for (int u = 0; u < U; ++u) slots.put(tmp_base + (int)dst_tmp, u, acc[u]);
    }
    if (dst_out != B200_DST_NONE) store((int)dst_out, acc);
  }
}

}  // namespace b200
