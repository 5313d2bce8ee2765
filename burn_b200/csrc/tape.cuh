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
// Design (B200): every thread owns U vectors of VEC consecutive elements.  The
// accumulator (result of the previous op) lives in registers; tape inputs and
// saved temporaries live in a thread-private shared-memory slot file laid out
// [slot][u][thread] as 16-byte words, so every LDS.128/STS.128 is conflict-free
// and no barrier is ever needed.  Opcode dispatch is warp-uniform (the tape
// sits in the kernel-parameter constant bank).
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int kTapeBlock = 256;  // threads per CTA in tape kernels
constexpr int kMaxDims = B200_MAX_RANK;

enum LoadMode : int32_t {
  kModeVec = 0,     // VEC consecutive elements, aligned vector access
  kModeBcast = 1,   // innermost stride 0: one element broadcast to all lanes
  kModeGather = 2   // VEC element accesses at offset + j*inner_stride
};

struct OperandDesc {
  void *ptr;
  int64_t strides[kMaxDims];  // in elements, collapsed dims, innermost last
  int32_t dtype;
  int32_t mode;
};

struct TapeParams {
  b200_tape_op ops[B200_MAX_TAPE_OPS];
  uint32_t scalars[B200_MAX_TAPE_SCALARS];
  OperandDesc in[B200_MAX_TAPE_INPUTS];
  OperandDesc out[B200_MAX_TAPE_OUTPUTS];
  FastDiv div[kMaxDims];  // div[d] divides by shape[d] (innermost measured in vectors)
  uint32_t shape[kMaxDims];
  int32_t n_ops, n_in, n_out, n_tmp, rank;
  uint32_t n_vec;  // number of VEC-wide vectors (numel / VEC)
};

// ----------------------------------------------------------------- scalar math
__device__ __forceinline__ float f_of(uint32_t u) { return __uint_as_float(u); }
__device__ __forceinline__ uint32_t u_of(float f) { return __float_as_uint(f); }

// erf evaluated the way the oracle does: libm::erf in f64, rounded to f32
// (crates/burn-ndarray/src/ops/tensor.rs:714-720).
__device__ __forceinline__ float erf_oracle(float x) { return (float)erf((double)x); }
__device__ __forceinline__ float tanh_oracle(float x) { return (float)tanh((double)x); }

// Python-style float modulo used by the reference for `remainder`
// (crates/burn-ndarray/src/ops/base.rs — `((x % rhs) + rhs) % rhs`).
__device__ __forceinline__ float rem_floor(float x, float y) {
  return fmodf(fmodf(x, y) + y, y);
}

__device__ __forceinline__ int32_t irem_floor(int32_t x, int32_t y) {
  if (y == 0) return 0;
  return ((x % y) + y) % y;
}

__device__ __forceinline__ float sign_f(float x) {
  // NaN stays NaN; burn `sign` → -1, 0, 1
  return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : x);
}

__device__ __forceinline__ float round_half_even(float x) { return rintf(x); }

static __device__ __noinline__ float pow_f(float a, float b) { return powf(a, b); }
static __device__ __noinline__ float slow_unary(int op, float x) {
  switch (op) {
    case B200_OP_SIN_F: return (float)sin((double)x);
    case B200_OP_COS_F: return (float)cos((double)x);
    case B200_OP_TAN_F: return (float)tan((double)x);
    case B200_OP_SINH_F: return (float)sinh((double)x);
    case B200_OP_COSH_F: return (float)cosh((double)x);
    case B200_OP_ASIN_F: return (float)asin((double)x);
    case B200_OP_ACOS_F: return (float)acos((double)x);
    case B200_OP_ATAN_F: return (float)atan((double)x);
    case B200_OP_ASINH_F: return (float)asinh((double)x);
    case B200_OP_ACOSH_F: return (float)acosh((double)x);
    case B200_OP_ATANH_F: return (float)atanh((double)x);
    default: return x;
  }
}

// ----------------------------------------------------------------- typed vector IO
// Loads VEC consecutive elements of `dtype` starting at element offset `off`
// as 32-bit lanes (f32 bits for float types, i32 for int/bool types).
template <int VEC>
__device__ __forceinline__ void load_vec(const void *base, int32_t dtype, int64_t off,
                                         uint32_t (&r)[VEC]) {
  if constexpr (VEC == 4) {
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
  } else {
#pragma unroll
    for (int j = 0; j < VEC; ++j) r[j] = 0;
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

template <int VEC>
__device__ __forceinline__ void store_vec(void *base, int32_t dtype, int64_t off,
                                          const uint32_t (&r)[VEC]) {
  if constexpr (VEC == 4) {
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
}

// Loads the VEC lanes of one operand for the vector whose collapsed
// coordinates are `coord` (innermost coordinate already in elements).
template <int VEC>
__device__ __forceinline__ void load_operand(const OperandDesc &d, int rank,
                                             const uint32_t (&coord)[kMaxDims],
                                             uint32_t (&r)[VEC]) {
  int64_t off = 0;
#pragma unroll
  for (int k = 0; k < kMaxDims; ++k)
    if (k < rank) off += (int64_t)coord[k] * d.strides[k];
  if (VEC == 4 && d.mode == kModeVec) {
    load_vec<VEC>(d.ptr, d.dtype, off, r);
  } else if (d.mode == kModeBcast) {
    const uint32_t v = load_one(d.ptr, d.dtype, off);
#pragma unroll
    for (int j = 0; j < VEC; ++j) r[j] = v;
  } else {
    const int64_t s = d.strides[rank - 1];
#pragma unroll
    for (int j = 0; j < VEC; ++j) r[j] = load_one(d.ptr, d.dtype, off + j * s);
  }
}

template <int VEC>
__device__ __forceinline__ void store_operand(const OperandDesc &d, int rank,
                                              const uint32_t (&coord)[kMaxDims],
                                              const uint32_t (&r)[VEC]) {
  int64_t off = 0;
#pragma unroll
  for (int k = 0; k < kMaxDims; ++k)
    if (k < rank) off += (int64_t)coord[k] * d.strides[k];
  if (VEC == 4 && d.mode == kModeVec) {
    store_vec<VEC>(d.ptr, d.dtype, off, r);
  } else {
    const int64_t s = d.strides[rank - 1];
#pragma unroll
    for (int j = 0; j < VEC; ++j) store_one(d.ptr, d.dtype, off + j * s, r[j]);
  }
}

// Decomposes vector index v into collapsed coordinates (innermost in elements).
template <int VEC>
__device__ __forceinline__ void vec_coords(const TapeParams &p, uint32_t v,
                                           uint32_t (&coord)[kMaxDims]) {
  uint32_t rest = v;
#pragma unroll
  for (int k = kMaxDims - 1; k >= 0; --k) {
    if (k < p.rank) {
      if (k == 0) {
        coord[k] = rest * (p.rank == 1 ? VEC : 1);
      } else {
        const uint32_t q = fd_div(rest, p.div[k]);
        const uint32_t r = rest - q * p.div[k].d;
        coord[k] = (k == p.rank - 1) ? r * VEC : r;
        rest = q;
      }
    } else {
      coord[k] = 0;
    }
  }
}

// ----------------------------------------------------------------- slot file
// Thread-private slots: word (slot, u) of thread t sits at
// smem[((slot*U + u) * kTapeBlock + t)] as a 16-byte word (VEC==4) or at
// 4-byte granularity (VEC==1).
template <int VEC, int U>
struct SlotFile {
  uint32_t *base;  // points at this thread's first word
  __device__ __forceinline__ void put(int slot, int u, const uint32_t (&r)[VEC]) const {
    if constexpr (VEC == 4) {
      *reinterpret_cast<uint4 *>(base + (size_t)(slot * U + u) * kTapeBlock * 4) =
          make_uint4(r[0], r[1], r[2], r[3]);
    } else {
      base[(size_t)(slot * U + u) * kTapeBlock] = r[0];
    }
  }
  __device__ __forceinline__ void get(int slot, int u, uint32_t (&r)[VEC]) const {
    if constexpr (VEC == 4) {
      const uint4 v = *reinterpret_cast<const uint4 *>(base + (size_t)(slot * U + u) * kTapeBlock * 4);
      r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
    } else {
      r[0] = base[(size_t)(slot * U + u) * kTapeBlock];
    }
  }
};

template <int VEC, int U>
__device__ __forceinline__ SlotFile<VEC, U> make_slots(uint32_t *smem, int tid) {
  SlotFile<VEC, U> s;
  s.base = smem + (VEC == 4 ? tid * 4 : tid);
  return s;
}

static inline size_t slot_file_bytes(int n_slots, int vec, int u) {
  return (size_t)n_slots * u * kTapeBlock * vec * 4;
}

// ----------------------------------------------------------------- interpreter
#define B200_UN(expr)                                   \
  _Pragma("unroll") for (int u = 0; u < U; ++u)         \
  _Pragma("unroll") for (int j = 0; j < VEC; ++j) {     \
    const float x = f_of(a[u][j]);                      \
    (void)x;                                            \
    acc[u][j] = u_of(expr);                             \
  }
#define B200_BIN(expr)                                  \
  _Pragma("unroll") for (int u = 0; u < U; ++u)         \
  _Pragma("unroll") for (int j = 0; j < VEC; ++j) {     \
    const float x = f_of(a[u][j]), y = f_of(b[u][j]);   \
    acc[u][j] = u_of(expr);                             \
  }
#define B200_CMP(expr)                                  \
  _Pragma("unroll") for (int u = 0; u < U; ++u)         \
  _Pragma("unroll") for (int j = 0; j < VEC; ++j) {     \
    const float x = f_of(a[u][j]), y = f_of(b[u][j]);   \
    (void)y;                                            \
    acc[u][j] = (expr) ? 1u : 0u;                       \
  }
#define B200_IUN(expr)                                  \
  _Pragma("unroll") for (int u = 0; u < U; ++u)         \
  _Pragma("unroll") for (int j = 0; j < VEC; ++j) {     \
    const int32_t x = (int32_t)a[u][j];                 \
    (void)x;                                            \
    acc[u][j] = (uint32_t)(expr);                       \
  }
#define B200_IBIN(expr)                                           \
  _Pragma("unroll") for (int u = 0; u < U; ++u)                   \
  _Pragma("unroll") for (int j = 0; j < VEC; ++j) {               \
    const int32_t x = (int32_t)a[u][j], y = (int32_t)b[u][j];     \
    (void)y;                                                      \
    acc[u][j] = (uint32_t)(expr);                                 \
  }

// Number of operands each opcode reads (1, 2 or 3).  Computed on the host and
// stored in b200_tape_op::pad[0] when a tape is copied into TapeParams.
static inline int op_arity(int op) {
  if (op == B200_OP_CLAMP_F || op == B200_OP_CLAMP_I || op == B200_OP_SELECT) return 3;
  switch (op) {
    case B200_OP_ADD_F: case B200_OP_SUB_F: case B200_OP_MUL_F: case B200_OP_DIV_F:
    case B200_OP_REM_F: case B200_OP_POW_F: case B200_OP_MIN_F: case B200_OP_MAX_F:
    case B200_OP_ATAN2_F:
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

template <int VEC, int U>
__device__ __forceinline__ void fetch_arg(uint8_t arg, int n_in, const uint32_t *scalars,
                                          const SlotFile<VEC, U> &slots,
                                          const uint32_t (&acc)[U][VEC],
                                          uint32_t (&dst)[U][VEC]) {
  const int kind = arg >> 6, idx = arg & 63;
  if (kind == 0) {
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < VEC; ++j) dst[u][j] = acc[u][j];
  } else if (kind == 3) {
    const uint32_t s = scalars[idx];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < VEC; ++j) dst[u][j] = s;
  } else {
    const int slot = (kind == 1) ? idx : n_in + idx;
#pragma unroll
    for (int u = 0; u < U; ++u) slots.get(slot, u, dst[u]);
  }
}

// Runs the whole tape.  `acc` must hold the initial accumulator on entry (used
// by GEMM/reduce write tapes through INPUT slots instead; elementwise tapes
// start with a MOV).  `store(out_index, acc)` is called for every op with a
// dst_out.
template <int VEC, int U, typename StoreFn>
__device__ __forceinline__ void run_tape(const TapeParams &p, const SlotFile<VEC, U> &slots,
                                         uint32_t (&acc)[U][VEC], StoreFn &&store) {
  for (int pc = 0; pc < p.n_ops; ++pc) {
    const b200_tape_op op = p.ops[pc];
    uint32_t a[U][VEC], b[U][VEC], c[U][VEC];
    const int ar = op.pad[0];
    fetch_arg<VEC, U>(op.a, p.n_in, p.scalars, slots, acc, a);
    if (ar >= 2) fetch_arg<VEC, U>(op.b, p.n_in, p.scalars, slots, acc, b);
    if (ar >= 3) fetch_arg<VEC, U>(op.c, p.n_in, p.scalars, slots, acc, c);
    switch (op.op) {
      case B200_OP_MOV: B200_UN(x) break;
      case B200_OP_ADD_F: B200_BIN(__fadd_rn(x, y)) break;
      case B200_OP_SUB_F: B200_BIN(__fsub_rn(x, y)) break;
      case B200_OP_MUL_F: B200_BIN(__fmul_rn(x, y)) break;
      case B200_OP_DIV_F: B200_BIN(__fdiv_rn(x, y)) break;
      case B200_OP_REM_F: B200_BIN(rem_floor(x, y)) break;
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
      case B200_OP_TANH_F: B200_UN(tanh_oracle(x)) break;
      case B200_OP_ERF_F: B200_UN(erf_oracle(x)) break;
      case B200_OP_FLOOR_F: B200_UN(floorf(x)) break;
      case B200_OP_CEIL_F: B200_UN(ceilf(x)) break;
      case B200_OP_ROUND_F: B200_UN(round_half_even(x)) break;
      case B200_OP_TRUNC_F: B200_UN(truncf(x)) break;
      case B200_OP_SIGN_F: B200_UN(sign_f(x)) break;
      case B200_OP_SIGMOID_F: B200_UN(__fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)))) break;
      case B200_OP_SIN_F: case B200_OP_COS_F: case B200_OP_TAN_F: case B200_OP_SINH_F:
      case B200_OP_COSH_F: case B200_OP_ASIN_F: case B200_OP_ACOS_F: case B200_OP_ATAN_F:
      case B200_OP_ASINH_F: case B200_OP_ACOSH_F: case B200_OP_ATANH_F:
        B200_UN(slow_unary(op.op, x)) break;
      case B200_OP_CLAMP_F:
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int j = 0; j < VEC; ++j) {
            const float x = f_of(a[u][j]), lo = f_of(b[u][j]), hi = f_of(c[u][j]);
            // Rust clamp semantics: NaN stays NaN
            acc[u][j] = u_of(x != x ? x : fminf(fmaxf(x, lo), hi));
          }
        break;
      case B200_OP_EQ_F: B200_CMP(x == y) break;
      case B200_OP_NE_F: B200_CMP(x != y) break;
      case B200_OP_LT_F: B200_CMP(x < y) break;
      case B200_OP_LE_F: B200_CMP(x <= y) break;
      case B200_OP_GT_F: B200_CMP(x > y) break;
      case B200_OP_GE_F: B200_CMP(x >= y) break;
      case B200_OP_ISNAN_F:
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int j = 0; j < VEC; ++j) acc[u][j] = (f_of(a[u][j]) != f_of(a[u][j])) ? 1u : 0u;
        break;
      case B200_OP_ISINF_F:
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int j = 0; j < VEC; ++j) acc[u][j] = isinf(f_of(a[u][j])) ? 1u : 0u;
        break;
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
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int j = 0; j < VEC; ++j)
            acc[u][j] = (uint32_t)min(max((int32_t)a[u][j], (int32_t)b[u][j]), (int32_t)c[u][j]);
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
      case B200_OP_SELECT:
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int j = 0; j < VEC; ++j) acc[u][j] = c[u][j] ? b[u][j] : a[u][j];
        break;
      case B200_OP_F2I: B200_UN(__int_as_float(__float2int_rz(x))) break;
      case B200_OP_I2F: B200_IUN(__float_as_int(__int2float_rn(x))) break;
      case B200_OP_B2F: B200_IUN(__float_as_int(x ? 1.0f : 0.0f)) break;
      case B200_OP_B2I: B200_IUN(x ? 1 : 0) break;
      case B200_OP_F2B: B200_UN(__int_as_float(x != 0.0f ? 1 : 0)) break;
      case B200_OP_I2B: B200_IUN(x != 0 ? 1 : 0) break;
      default: break;
    }
    if (op.dst_temp != B200_DST_NONE) {
#pragma unroll
      for (int u = 0; u < U; ++u) slots.put(p.n_in + op.dst_temp, u, acc[u]);
    }
    if (op.dst_out != B200_DST_NONE) store(op.dst_out, acc);
  }
}

}  // namespace b200
