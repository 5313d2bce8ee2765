// Multi-tensor Adam: one launch updates a whole flat bucket of parameters in place
// (SURVEY.md §8(f) row 3).  Per element, op for op and rounding for rounding, the sequence
// burn-optim records (crates/burn-optim/src/optim/adam.rs:149-210 AdaptiveMomentum::transform and
// :80-84 Adam::step):
//     m' = m*β1 + g*(1-β1)
//     v' = v*β2 + (g*g)*(1-β2)
//     u  = (m' * cf) / (sqrt(v') + eps_t)        cf = sqrt(1-β2^t)/(1-β1^t), eps_t = ε*sqrt(1-β2^t)
//     p' = p - u*lr
// cf and eps_t are read from device memory (`coef[0..1]`) so that a step captured in a CUDA graph
// replays with the current t.  Roofline: HBM, 28 B / element (read p,m,v,g; write p,m,v).
#include "common.cuh"

namespace b200 {
namespace optim {

struct AdamParams {
  float *p, *m, *v;
  const float *g;
  const float *coef;
  uint64_t n;
  float lr, b1, b2, omb1, omb2;
};

__device__ __forceinline__ void adam1(float &p, float &m, float &v, float g, const AdamParams &P, float cf, float eps_t) {
  m = __fadd_rn(__fmul_rn(m, P.b1), __fmul_rn(g, P.omb1));
  v = __fadd_rn(__fmul_rn(v, P.b2), __fmul_rn(__fmul_rn(g, g), P.omb2));
  const float u = __fdiv_rn(__fmul_rn(m, cf), __fadd_rn(__fsqrt_rn(v), eps_t));
  p = __fsub_rn(p, __fmul_rn(u, P.lr));
}

template <bool VEC>
__global__ void __launch_bounds__(256) adam_kernel(const AdamParams P) {
  const float cf = __ldg(P.coef), eps_t = __ldg(P.coef + 1);
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x, t0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if constexpr (VEC) {
    const uint64_t n4 = P.n >> 2;
    float4 *p4 = reinterpret_cast<float4 *>(P.p), *m4 = reinterpret_cast<float4 *>(P.m), *v4 = reinterpret_cast<float4 *>(P.v);
    const float4 *g4 = reinterpret_cast<const float4 *>(P.g);
    for (uint64_t i = t0; i < n4; i += stride) {
      float4 p = p4[i], m = m4[i], v = v4[i];
      const float4 g = __ldcs(g4 + i);
      adam1(p.x, m.x, v.x, g.x, P, cf, eps_t);
      adam1(p.y, m.y, v.y, g.y, P, cf, eps_t);
      adam1(p.z, m.z, v.z, g.z, P, cf, eps_t);
      adam1(p.w, m.w, v.w, g.w, P, cf, eps_t);
      p4[i] = p; m4[i] = m; v4[i] = v;
    }
    for (uint64_t i = (n4 << 2) + t0; i < P.n; i += stride) adam1(P.p[i], P.m[i], P.v[i], P.g[i], P, cf, eps_t);
  } else {
    for (uint64_t i = t0; i < P.n; i += stride) adam1(P.p[i], P.m[i], P.v[i], P.g[i], P, cf, eps_t);
  }
}

}  // namespace optim
}  // namespace b200

using namespace b200;

extern "C" int32_t b200_launch_adam(const b200_tensor *param, const b200_tensor *moment1, const b200_tensor *moment2,
                                    const b200_tensor *grad, const b200_tensor *coef, double lr, double beta1,
                                    double beta2, b200_stream s) {
  B200_REQUIRE(param && moment1 && moment2 && grad && coef, B200_ERR_INVALID, "null argument");
  int64_t n = 1;
  for (int d = 0; d < param->rank; ++d) n *= param->shape[d];
  for (const b200_tensor *t : {param, moment1, moment2, grad}) {
    B200_REQUIRE(t->dtype == B200_F32 && t->ptr, B200_ERR_UNSUPPORTED, "adam operands must be f32");
    B200_REQUIRE(is_contiguous(*t), B200_ERR_UNSUPPORTED, "adam operands must be contiguous");
    int64_t k = 1;
    for (int d = 0; d < t->rank; ++d) k *= t->shape[d];
    B200_REQUIRE(k == n, B200_ERR_SHAPE, "adam operands must have the same number of elements");
  }
  int64_t nc = 1;
  for (int d = 0; d < coef->rank; ++d) nc *= coef->shape[d];
  B200_REQUIRE(coef->dtype == B200_F32 && coef->ptr && nc >= 2 && is_contiguous(*coef), B200_ERR_INVALID,
               "coef must hold two contiguous f32 values");
  if (n == 0) return B200_OK;
  optim::AdamParams P;
  P.p = reinterpret_cast<float *>(param->ptr);
  P.m = reinterpret_cast<float *>(moment1->ptr);
  P.v = reinterpret_cast<float *>(moment2->ptr);
  P.g = reinterpret_cast<const float *>(grad->ptr);
  P.coef = reinterpret_cast<const float *>(coef->ptr);
  P.n = (uint64_t)n;
  P.lr = (float)lr;
  P.b1 = (float)beta1;
  P.b2 = (float)beta2;
  P.omb1 = 1.0f - (float)beta1;   // f32 arithmetic, like `1.0 - self.beta_1` on the reference's f32 fields
  P.omb2 = 1.0f - (float)beta2;
  const bool vec = (((uintptr_t)P.p | (uintptr_t)P.m | (uintptr_t)P.v | (uintptr_t)P.g) & 15) == 0;
  const uint64_t work = vec ? (P.n + 3) / 4 : P.n;
  const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((work + 255) / 256, (uint64_t)sm_count() * 8));
  if (vec) optim::adam_kernel<true><<<grid, 256, 0, resolve_stream(s)>>>(P);
  else optim::adam_kernel<false><<<grid, 256, 0, resolve_stream(s)>>>(P);
  B200_LAUNCH_CHECK();
  return B200_OK;
}
