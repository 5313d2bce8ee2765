// Scalar math of the op-tape ISA, shared by the interpreter (tape.cuh), the GEMM epilogue and the
// NVRTC-specialised kernels (jit.cu compiles this header at run time, so it depends on nothing but
// burn_b200.h's opcode enum and CUDA built-ins).
#pragma once
#include "erf_table.inc"
#include "tanh_table.inc"

namespace b200 {
// Internal opcodes = public opcodes + a few compiler-generated ones.
enum : int {
  kOpLoad = B200_OP_COUNT,  // acc = B
  kOpSave,                  // no-op carrying a dst_temp (acc saved to a temp)
  kOpDivScalar,             // acc = acc / B, B a scalar: exact Markstein sequence
  kOpMulAdd,                // acc = RN(RN(acc * B) + C)   (two roundings, never an FMA)
  kOpGelu,                  // acc = gelu(B): the reference's 5-op chain executed in one dispatch
  kOpSquare,                // acc = acc*acc          powf_scalar(x, 2): exactly the correctly rounded power
  kOpCube,                  // acc = (acc*acc)*acc    powf_scalar(x, 3) = powi_scalar(x, 3): within 1 ulp of it (libm powf: 1-4 ulp)
  kIOpCount
};


// ----------------------------------------------------------------- scalar math
__device__ __forceinline__ float f_of(uint32_t u) { return __uint_as_float(u); }
__device__ __forceinline__ uint32_t u_of(float f) { return __float_as_uint(f); }

// tanh.  The oracle computes tanh in f64 and rounds to f32 (crates/burn-ndarray/src/ops/tensor.rs:620-626).  Round 1
// did exactly that on the device; ~150 FP64 instructions per element made the reference's gelu_backward (one tanh per
// element of a [tokens, d_ff] tensor) run at a third of the HBM roofline.  This is an f32 routine built like erf_f32
// (scripts/fit_tanh.py): within 1 ulp of the correctly rounded value everywhere, equal to it for > 98 % of inputs.
//   |x| < 0.5          : fma(x, t*Q(t), x), t = x^2
//   0.5 <= |x| < 9.125 : 69 table intervals of width 1/8: H_i + (d*Q_i(d) + L_i), d = |x| - c_i exact
//   |x| >= 9.125       : +-1
__device__ __forceinline__ float tanh_f32(float x) {
  const float ax = fabsf(x);
  const float t = __fmul_rn(ax, ax);
  float qa = B200_TANH_Q5;
  qa = __fmaf_rn(qa, t, B200_TANH_Q4);
  qa = __fmaf_rn(qa, t, B200_TANH_Q3);
  qa = __fmaf_rn(qa, t, B200_TANH_Q2);
  qa = __fmaf_rn(qa, t, B200_TANH_Q1);
  qa = __fmaf_rn(qa, t, B200_TANH_Q0);
  const float ra = __fmaf_rn(ax, __fmul_rn(t, qa), ax);
  const int i = min(max(__float2int_rz(__fmul_rn(__fsub_rn(ax, 0.5f), 8.0f)), 0), 68);
  const float c = __fmaf_rn((float)i, 0.125f, 0.5625f);
  const float d = __fsub_rn(ax, c);
  const float4 r0 = __ldg(&kTanhB[2 * i]), r1 = __ldg(&kTanhB[2 * i + 1]);
  float qb = r1.z;
  qb = __fmaf_rn(qb, d, r1.y);
  qb = __fmaf_rn(qb, d, r1.x);
  qb = __fmaf_rn(qb, d, r0.w);
  qb = __fmaf_rn(qb, d, r0.z);
  const float rb = __fadd_rn(r0.x, __fmaf_rn(d, qb, r0.y));
  float r = ax < 0.5f ? ra : rb;
  r = ax >= 9.125f ? 1.0f : r;
  return copysignf(r, x);
}

// erf.  The oracle computes libm::erf in f64 and rounds to f32
// (crates/burn-ndarray/src/ops/tensor.rs:714-720); FP64 runs at half rate on B200
// and would cap a fused gelu chain below the HBM roofline, so this is an f32
// routine designed (scripts/fit_erf.py) to land within 1 ulp of that correctly
// rounded value — equal to it for ~99% of inputs:
//   |x| < 0.75      : x*K_hi + x*(K_lo + t*Q(t)), t = x^2 (single final rounding), Q of degree 6
//   0.75 <= |x| < 4 : table intervals of width 1/8: H_i + (d*Q_i(d) + L_i), d = |x| - c_i exact
//   |x| >= 4        : +-1 (erfc(4) < half an ulp of 1)
// The table region costs more than half of the instructions; the lanes that execute together skip it when none of
// them needs it (a gelu of inputs within +-1 — the fuse-on-read reduce of the benchmark — never does).
__device__ __forceinline__ float erf_f32(float x) {
  const float ax = fabsf(x);
  // region A
  const float t = __fmul_rn(ax, ax);
  float qa = B200_ERF_Q6;
  qa = __fmaf_rn(qa, t, B200_ERF_Q5);
  qa = __fmaf_rn(qa, t, B200_ERF_Q4);
  qa = __fmaf_rn(qa, t, B200_ERF_Q3);
  qa = __fmaf_rn(qa, t, B200_ERF_Q2);
  qa = __fmaf_rn(qa, t, B200_ERF_Q1);
  qa = __fmaf_rn(qa, t, B200_ERF_Q0);
  const float e = __fmaf_rn(t, qa, B200_ERF_K_LO);
  float r = __fmaf_rn(ax, B200_ERF_K_HI, __fmul_rn(ax, e));
  const bool outer = !(ax < 0.75f);                      // NaN takes the table path (row 0 -> NaN)
  if (__any_sync(__activemask(), outer)) {
    // region B (index clamped so the table read is always in range)
    const int i = min(max(__float2int_rz(__fmul_rn(__fsub_rn(ax, 0.5f), 8.0f)), 0), 27);
    const float c = __fmaf_rn((float)i, 0.125f, 0.5625f);
    const float d = __fsub_rn(ax, c);
    const float4 r0 = __ldg(&kErfB[2 * i]), r1 = __ldg(&kErfB[2 * i + 1]);
    float qb = r1.z;
    qb = __fmaf_rn(qb, d, r1.y);
    qb = __fmaf_rn(qb, d, r1.x);
    qb = __fmaf_rn(qb, d, r0.w);
    qb = __fmaf_rn(qb, d, r0.z);
    const float rb = __fadd_rn(r0.x, __fmaf_rn(d, qb, r0.y));
    r = outer ? rb : r;
    r = ax >= 4.0f ? 1.0f : r;
  }
  return copysignf(r, x);
}

// Python-style float modulo used by the reference for `remainder`
// (crates/burn-ndarray/src/ops/base.rs — `((x % rhs) + rhs) % rhs`).
__device__ __forceinline__ float rem_floor(float x, float y) {
  return fmodf(fmodf(x, y) + y, y);
}

// tensor-tensor `remainder` of the reference: a - b*floor(a/b) in f64, then rounded (base.rs:909-922)
__device__ __forceinline__ float rem_tensor(float x, float y) {
  const double a = (double)x, b = (double)y;
  return (float)__dsub_rn(a, __dmul_rn(b, floor(__ddiv_rn(a, b))));   // no FMA contraction: Rust rounds the product
}

__device__ __forceinline__ int32_t irem_floor(int32_t x, int32_t y) {
  if (y == 0) return 0;
  return ((x % y) + y) % y;
}

__device__ __forceinline__ float sign_f(float x) {
  return (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : x);  // NaN stays NaN
}

// x / y for a launch-constant y with correctly rounded rinv = RN(1/y): Markstein's
// sequence gives the correctly rounded quotient when nothing under/overflows; the
// guarded range falls back to the IEEE division.  (Host only emits kOpDivScalar
// for normal y whose significand is not all ones.)
__device__ __forceinline__ float div_scalar_fast(float x, float y, float rinv) {
  const float q = __fmul_rn(x, rinv);
  const float r = __fmaf_rn(-y, q, x);
  // r == 0: q is already the exact quotient — returning it also keeps the sign of a zero
  // quotient (fma(+0, rinv, -0) would give +0 for x = -0)
  return r == 0.0f ? q : __fmaf_rn(r, rinv, q);
}
// True when Markstein's sequence is exact for x (no intermediate under/overflow);
// zero is fine (0/y = 0).
__device__ __forceinline__ bool div_scalar_safe(float x) {
  const float ax = fabsf(x);
  return (ax > 1e-25f && ax < 1e25f) || ax == 0.0f;
}
__device__ __forceinline__ float div_scalar_exact(float x, float y, float rinv) {
  return div_scalar_safe(x) ? div_scalar_fast(x, y, rinv) : __fdiv_rn(x, y);
}

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

}  // namespace b200
