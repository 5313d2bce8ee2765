// One op of the accumulator ISA (tape.cuh) on ONE element, with the opcode as a compile-time
// constant: the building block of the NVRTC-specialised kernels jit.cu generates from a compiled
// tape.  Every case performs exactly the operation of the interpreter's switch in tape.cuh run_tape
// (same intrinsics, same rounding points), so a specialised kernel and the interpreter agree bit
// for bit; tests/test_jit_gpu.py holds them to that.
#pragma once
#include "tape_math.cuh"

namespace b200 {

// Kernel parameter block of a specialised kernel (operand base pointers are already offset).
struct JitParams {
  const void *in[B200_MAX_TAPE_INPUTS];
  void *out[B200_MAX_TAPE_OUTPUTS];
  uint32_t scalars[B200_MAX_TAPE_SCALARS];
  uint32_t n_vec;
};

template <int OPC>
__device__ __forceinline__ uint32_t eval_op(uint32_t a, uint32_t b, uint32_t c) {
  const float x = f_of(a), y = f_of(b);
  const int32_t xi = (int32_t)a, yi = (int32_t)b;
  (void)x; (void)y; (void)xi; (void)yi; (void)c;
  switch (OPC) {
    case kOpLoad: return b;
    case kOpSave: return a;
    case B200_OP_MOV: return b;
    case B200_OP_ADD_F: return u_of(__fadd_rn(x, y));
    case B200_OP_SUB_F: return u_of(__fsub_rn(x, y));
    case B200_OP_MUL_F: return u_of(__fmul_rn(x, y));
    case B200_OP_DIV_F: return u_of(__fdiv_rn(x, y));
    case kOpDivScalar: return u_of(div_scalar_exact(x, y, __frcp_rn(y)));
    case kOpMulAdd: return u_of(__fadd_rn(__fmul_rn(x, y), f_of(c)));
    case kOpSquare: return u_of(__fmul_rn(x, x));
    case kOpCube: return u_of(__fmul_rn(__fmul_rn(x, x), x));
    case kOpGelu: {  // gelu of B (the generator passes the accumulator as B when the op has none)
      const float s2 = 1.41421353816986083984375f, rinv = 0.707106769084930419921875f;
      const float e = erf_f32(div_scalar_exact(y, s2, rinv));
      return u_of(__fmul_rn(__fmul_rn(y, __fadd_rn(e, 1.0f)), 0.5f));
    }
    case B200_OP_REM_F: return u_of(rem_floor(x, y));
    case B200_OP_REMT_F: return u_of(rem_tensor(x, y));
    case B200_OP_POW_F: return u_of(pow_f(x, y));
    case B200_OP_MIN_F: return u_of((x != x || y != y) ? __int_as_float(0x7fc00000) : fminf(x, y));
    case B200_OP_MAX_F: return u_of((x != x || y != y) ? __int_as_float(0x7fc00000) : fmaxf(x, y));
    case B200_OP_ATAN2_F: return u_of((float)atan2((double)x, (double)y));
    case B200_OP_NEG_F: return u_of(-x);
    case B200_OP_ABS_F: return u_of(fabsf(x));
    case B200_OP_EXP_F: return u_of(expf(x));
    case B200_OP_LOG_F: return u_of(logf(x));
    case B200_OP_LOG1P_F: return u_of(log1pf(x));
    case B200_OP_SQRT_F: return u_of(__fsqrt_rn(x));
    case B200_OP_RECIP_F: return u_of(__fdiv_rn(1.0f, x));
    case B200_OP_TANH_F: return u_of(tanh_f32(x));
    case B200_OP_ERF_F: return u_of(erf_f32(x));
    case B200_OP_FLOOR_F: return u_of(floorf(x));
    case B200_OP_CEIL_F: return u_of(ceilf(x));
    case B200_OP_ROUND_F: return u_of(rintf(x));
    case B200_OP_TRUNC_F: return u_of(truncf(x));
    case B200_OP_SIGN_F: return u_of(sign_f(x));
    case B200_OP_SIGMOID_F: return u_of(__fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))));
    case B200_OP_SIN_F: case B200_OP_COS_F: case B200_OP_TAN_F: case B200_OP_SINH_F:
    case B200_OP_COSH_F: case B200_OP_ASIN_F: case B200_OP_ACOS_F: case B200_OP_ATAN_F:
    case B200_OP_ASINH_F: case B200_OP_ACOSH_F: case B200_OP_ATANH_F:
      return u_of(slow_unary(OPC, x));
    case B200_OP_CLAMP_F: {
      const float lo = y, hi = f_of(c);
      return u_of(x != x ? x : fminf(fmaxf(x, lo), hi));
    }
    case B200_OP_EQ_F: return x == y ? 1u : 0u;
    case B200_OP_NE_F: return x != y ? 1u : 0u;
    case B200_OP_LT_F: return x < y ? 1u : 0u;
    case B200_OP_LE_F: return x <= y ? 1u : 0u;
    case B200_OP_GT_F: return x > y ? 1u : 0u;
    case B200_OP_GE_F: return x >= y ? 1u : 0u;
    case B200_OP_ISNAN_F: return x != x ? 1u : 0u;
    case B200_OP_ISINF_F: return isinf(x) ? 1u : 0u;
    case B200_OP_ADD_I: return (uint32_t)(xi + yi);
    case B200_OP_SUB_I: return (uint32_t)(xi - yi);
    case B200_OP_MUL_I: return (uint32_t)(xi * yi);
    case B200_OP_DIV_I: return (uint32_t)(yi == 0 ? 0 : xi / yi);
    case B200_OP_REM_I: return (uint32_t)irem_floor(xi, yi);
    case B200_OP_MIN_I: return (uint32_t)min(xi, yi);
    case B200_OP_MAX_I: return (uint32_t)max(xi, yi);
    case B200_OP_NEG_I: return (uint32_t)(-xi);
    case B200_OP_ABS_I: return (uint32_t)abs(xi);
    case B200_OP_SIGN_I: return (uint32_t)((xi > 0) - (xi < 0));
    case B200_OP_AND_I: return (uint32_t)(xi & yi);
    case B200_OP_OR_I: return (uint32_t)(xi | yi);
    case B200_OP_XOR_I: return (uint32_t)(xi ^ yi);
    case B200_OP_NOT_I: return (uint32_t)(~xi);
    case B200_OP_SHL_I: return (uint32_t)(xi << (yi & 31));
    case B200_OP_SHR_I: return (uint32_t)(xi >> (yi & 31));
    case B200_OP_CLAMP_I: return (uint32_t)min(max(xi, yi), (int32_t)c);
    case B200_OP_EQ_I: return xi == yi ? 1u : 0u;
    case B200_OP_NE_I: return xi != yi ? 1u : 0u;
    case B200_OP_LT_I: return xi < yi ? 1u : 0u;
    case B200_OP_LE_I: return xi <= yi ? 1u : 0u;
    case B200_OP_GT_I: return xi > yi ? 1u : 0u;
    case B200_OP_GE_I: return xi >= yi ? 1u : 0u;
    case B200_OP_AND_B: return (uint32_t)((xi != 0) & (yi != 0));
    case B200_OP_OR_B: return (uint32_t)((xi != 0) | (yi != 0));
    case B200_OP_XOR_B: return (uint32_t)((xi != 0) ^ (yi != 0));
    case B200_OP_NOT_B: return xi == 0 ? 1u : 0u;
    case B200_OP_SELECT: return c ? b : a;
    case B200_OP_F2I: return (uint32_t)__float2int_rz(x);
    case B200_OP_I2F: return u_of(__int2float_rn(xi));
    case B200_OP_B2F: return u_of(xi ? 1.0f : 0.0f);
    case B200_OP_B2I: return xi ? 1u : 0u;
    case B200_OP_F2B: return x != 0.0f ? 1u : 0u;
    case B200_OP_I2B: return xi != 0 ? 1u : 0u;
    default: return a;
  }
}

// ---- typed 4-element vector IO of the specialised kernels (subset of tape.cuh load_vec4/store_vec4)
// Half-precision conversions as PTX (NVRTC sees no cuda_fp16.h): the instructions __half22float2,
// __floats2half2_rn and __floats2bfloat162_rn compile to, so both execution modes round identically.
__device__ __forceinline__ void jit_h2_to_f2(uint32_t h2, uint32_t &lo, uint32_t &hi) {
  asm("{ .reg .b16 l, h; mov.b32 {l, h}, %2; cvt.f32.f16 %0, l; cvt.f32.f16 %1, h; }" : "=r"(lo), "=r"(hi) : "r"(h2));
}
__device__ __forceinline__ uint32_t jit_h1_to_f(unsigned short h) {
  uint32_t f;
  asm("{ .reg .b16 l; mov.b16 l, %1; cvt.f32.f16 %0, l; }" : "=r"(f) : "h"(h));
  return f;
}
__device__ __forceinline__ uint32_t jit_f2_to_h2(uint32_t lo, uint32_t hi) {
  uint32_t d;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(f_of(hi)), "f"(f_of(lo)));
  return d;
}
__device__ __forceinline__ uint32_t jit_f2_to_bf2(uint32_t lo, uint32_t hi) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(f_of(hi)), "f"(f_of(lo)));
  return d;
}
template <int DT>
__device__ __forceinline__ void jit_ld4(const void *base, uint32_t v, uint32_t (&r)[4]) {
  if (DT == B200_F32 || DT == B200_I32) {
    const uint4 q = __ldcs(reinterpret_cast<const uint4 *>(base) + v);
    r[0] = q.x; r[1] = q.y; r[2] = q.z; r[3] = q.w;
  } else if (DT == B200_BF16) {
    const uint2 q = __ldcs(reinterpret_cast<const uint2 *>(base) + v);
    r[0] = q.x << 16; r[1] = q.x & 0xFFFF0000u; r[2] = q.y << 16; r[3] = q.y & 0xFFFF0000u;
  } else if (DT == B200_F16) {
    const uint2 q = __ldcs(reinterpret_cast<const uint2 *>(base) + v);
    jit_h2_to_f2(q.x, r[0], r[1]);
    jit_h2_to_f2(q.y, r[2], r[3]);
  } else {  // BOOL / U8
    const uint32_t q = __ldcs(reinterpret_cast<const uint32_t *>(base) + v);
    r[0] = q & 0xFFu; r[1] = (q >> 8) & 0xFFu; r[2] = (q >> 16) & 0xFFu; r[3] = q >> 24;
  }
}
template <int DT>
__device__ __forceinline__ uint32_t jit_ld1(const void *base) {
  if (DT == B200_F32 || DT == B200_I32) return __ldg(reinterpret_cast<const uint32_t *>(base));
  if (DT == B200_BF16) return (uint32_t)__ldg(reinterpret_cast<const unsigned short *>(base)) << 16;
  if (DT == B200_F16) return jit_h1_to_f(__ldg(reinterpret_cast<const unsigned short *>(base)));
  return (uint32_t)__ldg(reinterpret_cast<const unsigned char *>(base));
}
template <int DT>
__device__ __forceinline__ uint32_t jit_ld1_at(const void *base, uint32_t off) {
  if (DT == B200_F32 || DT == B200_I32) return __ldg(reinterpret_cast<const uint32_t *>(base) + off);
  if (DT == B200_BF16) return (uint32_t)__ldg(reinterpret_cast<const unsigned short *>(base) + off) << 16;
  if (DT == B200_F16) return jit_h1_to_f(__ldg(reinterpret_cast<const unsigned short *>(base) + off));
  return (uint32_t)__ldg(reinterpret_cast<const unsigned char *>(base) + off);
}
template <int DT>
__device__ __forceinline__ void jit_st4(void *base, uint32_t v, const uint32_t (&r)[4]) {
  if (DT == B200_F32 || DT == B200_I32) {
    __stcs(reinterpret_cast<uint4 *>(base) + v, make_uint4(r[0], r[1], r[2], r[3]));
  } else if (DT == B200_BF16) {
    __stcs(reinterpret_cast<uint2 *>(base) + v, make_uint2(jit_f2_to_bf2(r[0], r[1]), jit_f2_to_bf2(r[2], r[3])));
  } else if (DT == B200_F16) {
    __stcs(reinterpret_cast<uint2 *>(base) + v, make_uint2(jit_f2_to_h2(r[0], r[1]), jit_f2_to_h2(r[2], r[3])));
  } else {  // BOOL / U8
    __stcs(reinterpret_cast<uint32_t *>(base) + v,
           (r[0] & 0xFFu) | ((r[1] & 0xFFu) << 8) | ((r[2] & 0xFFu) << 16) | (r[3] << 24));
  }
}

}  // namespace b200
