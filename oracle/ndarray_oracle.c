/*
 * ndarray_oracle.c — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the burn-ndarray semantics for the burn-b200 hot path.  It
 * exists so the CUDA kernels can be checked for parity; it is never linked,
 * loaded or called by the product (burn_b200/), only by tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke().
 *
 * Parity status: PINNED against the reference's own golden vectors — every
 * function here is checked in tests/test_oracle_golden.py against literal
 * expected values transliterated from crates/burn-backend-tests/tests/tensor/**
 * (file:line cited per fixture in tests/golden/burn_backend_tests.json).  The
 * reference itself (Rust; needs cargo + the un-vendored crates ndarray 0.17.2,
 * matrixmultiply 0.3.11, libm 0.2.16, macerator 0.3.4) cannot be built here.
 *
 * Third-party arithmetic restated from its published algorithm:
 *  - ndarray 0.17.2 `numeric_util::unrolled_fold` (8 partial sums combined as
 *    (p0+p4)+(p1+p5)+(p2+p6)+(p3+p7), then the <8 tail sequentially) used by
 *    `ArrayBase::sum`, and `sum_axis` (lane-wise unrolled_fold when the reduced
 *    axis is the min-stride axis, otherwise `res = res + subview` for each index
 *    of the axis).  Call sites: crates/burn-ndarray/src/ops/base.rs:940-943,
 *    crates/burn-ndarray/src/ops/macros.rs:45-60.
 *  - matrixmultiply 0.3.11 `sgemm`: C += A_blk·B_blk over K blocks of KC=256,
 *    each block accumulated from zero in k order with fused multiply-add.
 *    Call site: crates/burn-ndarray/src/ops/matmul.rs:55-61.
 *  - libm 0.2.16 `erf` (f64): restated with the C library's erf(); both are
 *    faithful (<1 ulp in f64) and the result is rounded to f32
 *    (crates/burn-ndarray/src/ops/tensor.rs:714-720).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define API __attribute__((visibility("default")))

/* ---- elementwise -------------------------------------------------------- */
/* op codes are private to the oracle (oracle/oracle.py mirrors them) */
enum { O_ADD = 0, O_SUB, O_MUL, O_DIV, O_REM, O_POW, O_MIN, O_MAX, O_REMT };

static inline float bin_f32(int op, float x, float y) {
  switch (op) {
    case O_ADD: return x + y;
    case O_SUB: return x - y;
    case O_MUL: return x * y;
    case O_DIV: return x / y;
    /* crates/burn-ndarray/src/ops/base.rs:924-930 remainder_scalar: ((x % y) + y) % y */
    case O_REM: return fmodf(fmodf(x, y) + y, y);
    /* crates/burn-ndarray/src/ops/base.rs:909-922 remainder (tensor-tensor): a - b*floor(a/b) in f64 */
    case O_REMT: { const double a = (double)x, b = (double)y; return (float)(a - b * floor(a / b)); }
    case O_POW: return powf(x, y);
    case O_MIN: return (x != x || y != y) ? NAN : (x < y ? x : y);
    case O_MAX: return (x != x || y != y) ? NAN : (x > y ? x : y);
    default: return NAN;
  }
}

/* crates/burn-ndarray/src/ops/base.rs:807+ (add/sub/mul/div on same-shape arrays) */
API void o_binary_f32(int op, const float *a, const float *b, float *out, size_t n) {
  for (size_t i = 0; i < n; ++i) out[i] = bin_f32(op, a[i], b[i]);
}

API void o_binary_scalar_f32(int op, const float *a, float s, float *out, size_t n) {
  for (size_t i = 0; i < n; ++i) out[i] = bin_f32(op, a[i], s);
}

enum {
  U_EXP = 0, U_LOG, U_LOG1P, U_SQRT, U_TANH, U_ERF, U_SIN, U_COS, U_TAN, U_RECIP, U_ABS, U_NEG,
  U_FLOOR, U_CEIL, U_ROUND, U_TRUNC, U_SINH, U_COSH, U_ASIN, U_ACOS, U_ATAN, U_ASINH, U_ACOSH,
  U_ATANH, U_SIGN
};

/* crates/burn-ndarray/src/ops/tensor.rs:515-720 — exp/log/log1p/sqrt in f32
 * (element.rs:100-140: self.exp(), self.ln(), …), the rest evaluated in f64 and
 * rounded to f32. */
API void o_unary_f32(int op, const float *a, float *out, size_t n) {
  for (size_t i = 0; i < n; ++i) {
    const float x = a[i];
    float r;
    switch (op) {
      case U_EXP: r = expf(x); break;
      case U_LOG: r = logf(x); break;
      case U_LOG1P: r = log1pf(x); break;
      case U_SQRT: r = sqrtf(x); break;
      case U_TANH: r = (float)tanh((double)x); break;
      case U_ERF: r = (float)erf((double)x); break;
      case U_SIN: r = (float)sin((double)x); break;
      case U_COS: r = (float)cos((double)x); break;
      case U_TAN: r = (float)tan((double)x); break;
      case U_RECIP: r = 1.0f / x; break;
      case U_ABS: r = fabsf(x); break;
      case U_NEG: r = -x; break;
      case U_FLOOR: r = (float)floor((double)x); break;
      case U_CEIL: r = (float)ceil((double)x); break;
      case U_ROUND: r = (float)rint((double)x); break; /* round half to even */
      case U_TRUNC: r = (float)trunc((double)x); break;
      case U_SINH: r = (float)sinh((double)x); break;
      case U_COSH: r = (float)cosh((double)x); break;
      case U_ASIN: r = (float)asin((double)x); break;
      case U_ACOS: r = (float)acos((double)x); break;
      case U_ATAN: r = (float)atan((double)x); break;
      case U_ASINH: r = (float)asinh((double)x); break;
      case U_ACOSH: r = (float)acosh((double)x); break;
      case U_ATANH: r = (float)atanh((double)x); break;
      case U_SIGN: r = (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : x); break;
      default: r = NAN;
    }
    out[i] = r;
  }
}

/* mask_fill / mask_where — crates/burn-ndarray/src/ops/base.rs:78-104 */
API void o_mask_fill_f32(const float *a, const uint8_t *mask, float value, float *out, size_t n) {
  for (size_t i = 0; i < n; ++i) out[i] = mask[i] ? value : a[i];
}
API void o_mask_where_f32(const float *a, const uint8_t *mask, const float *src, float *out, size_t n) {
  for (size_t i = 0; i < n; ++i) out[i] = mask[i] ? src[i] : a[i];
}

/* ---- reductions --------------------------------------------------------- */
/* ndarray numeric_util::unrolled_fold with f = add, init = 0 */
static float unrolled_sum_f32(const float *xs, size_t n) {
  float acc = 0.f;
  float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f, p4 = 0.f, p5 = 0.f, p6 = 0.f, p7 = 0.f;
  while (n >= 8) {
    p0 += xs[0]; p1 += xs[1]; p2 += xs[2]; p3 += xs[3];
    p4 += xs[4]; p5 += xs[5]; p6 += xs[6]; p7 += xs[7];
    xs += 8; n -= 8;
  }
  acc = acc + (p0 + p4);
  acc = acc + (p1 + p5);
  acc = acc + (p2 + p6);
  acc = acc + (p3 + p7);
  for (size_t i = 0; i < n; ++i) acc = acc + xs[i];
  return acc;
}

/* float_sum → sum_view → ArrayView::sum (contiguous) */
API float o_sum_f32(const float *x, size_t n) { return unrolled_sum_f32(x, n); }

/* float_sum_dim → sum_axis on a C-contiguous [outer, R, inner] view */
API void o_sum_axis_f32(const float *x, size_t outer, size_t R, size_t inner, float *out) {
  if (inner == 1) {
    for (size_t o = 0; o < outer; ++o) out[o] = unrolled_sum_f32(x + o * R, R);
    return;
  }
  for (size_t o = 0; o < outer; ++o) {
    float *dst = out + o * inner;
    for (size_t c = 0; c < inner; ++c) dst[c] = 0.f;
    for (size_t r = 0; r < R; ++r) {
      const float *src = x + (o * R + r) * inner;
      for (size_t c = 0; c < inner; ++c) dst[c] = dst[c] + src[c];
    }
  }
}

/* float_mean_dim → mean_axis = sum_axis / len */
API void o_mean_axis_f32(const float *x, size_t outer, size_t R, size_t inner, float *out) {
  o_sum_axis_f32(x, outer, R, inner, out);
  const float n = (float)R;
  for (size_t i = 0; i < outer * inner; ++i) out[i] = out[i] / n;
}

API void o_prod_axis_f32(const float *x, size_t outer, size_t R, size_t inner, float *out) {
  for (size_t o = 0; o < outer; ++o)
    for (size_t c = 0; c < inner; ++c) {
      float acc = 1.f;
      for (size_t r = 0; r < R; ++r) acc = acc * x[(o * R + r) * inner + c];
      out[o * inner + c] = acc;
    }
}

/* arg_view — crates/burn-ndarray/src/ops/base.rs:1715-1757 */
API void o_arg_f32(const float *x, size_t outer, size_t R, size_t inner, int is_min, int64_t *out) {
  for (size_t o = 0; o < outer; ++o)
    for (size_t c = 0; c < inner; ++c) {
      float best = x[(o * R) * inner + c];
      size_t bi = 0;
      for (size_t r = 0; r < R; ++r) {
        const float e = x[(o * R + r) * inner + c];
        const int acc_nan = best != best, e_nan = e != e;
        const int take = !acc_nan && (e_nan || (is_min ? (e < best) : (e > best)));
        if (take) { best = e; bi = r; }
      }
      out[o * inner + c] = (int64_t)bi;
    }
}

API void o_arg_i64(const int64_t *x, size_t outer, size_t R, size_t inner, int is_min, int64_t *out) {
  for (size_t o = 0; o < outer; ++o)
    for (size_t c = 0; c < inner; ++c) {
      int64_t best = x[(o * R) * inner + c];
      size_t bi = 0;
      for (size_t r = 0; r < R; ++r) {
        const int64_t e = x[(o * R + r) * inner + c];
        if (is_min ? (e < best) : (e > best)) { best = e; bi = r; }
      }
      out[o * inner + c] = (int64_t)bi;
    }
}

/* ---- matmul ------------------------------------------------------------- */
/* One batch of general_mat_mul(1, A, B, 0, C): row-major with explicit strides
 * (crates/burn-ndarray/src/ops/matmul.rs:55-61).  KC-blocked f32 FMA
 * accumulation as matrixmultiply::sgemm does. */
API void o_sgemm(const float *a, int64_t a_rs, int64_t a_cs, const float *b, int64_t b_rs,
                 int64_t b_cs, float *c, int64_t M, int64_t N, int64_t K) {
  const int64_t KC = 256;
  float *blk = (float *)malloc(sizeof(float) * (size_t)N);
  for (int64_t i = 0; i < M; ++i) {
    float *crow = c + i * N;
    for (int64_t j = 0; j < N; ++j) crow[j] = 0.f;
    for (int64_t k0 = 0; k0 < K; k0 += KC) {
      const int64_t k1 = k0 + KC < K ? k0 + KC : K;
      for (int64_t j = 0; j < N; ++j) blk[j] = 0.f;
      for (int64_t k = k0; k < k1; ++k) {
        const float av = a[i * a_rs + k * a_cs];
        const float *brow = b + k * b_rs;
        if (b_cs == 1) {
          for (int64_t j = 0; j < N; ++j) blk[j] = fmaf(av, brow[j], blk[j]);
        } else {
          for (int64_t j = 0; j < N; ++j) blk[j] = fmaf(av, brow[j * b_cs], blk[j]);
        }
      }
      for (int64_t j = 0; j < N; ++j) crow[j] = crow[j] + blk[j];
    }
  }
  free(blk);
}

/* ---- indexing ----------------------------------------------------------- */
/* gather along `dim` of a C-contiguous tensor viewed as [outer, D, inner];
 * indices/out are [outer, DI, inner].  crates/burn-ndarray/src/ops/base.rs:106-138 */
API void o_gather_f32(const float *t, const int64_t *idx, float *out, size_t outer, size_t D,
                      size_t DI, size_t inner) {
  for (size_t o = 0; o < outer; ++o)
    for (size_t i = 0; i < DI; ++i)
      for (size_t c = 0; c < inner; ++c) {
        const size_t k = (o * DI + i) * inner + c;
        out[k] = t[(o * D + (size_t)idx[k]) * inner + c];
      }
}

/* scatter-add: t[.., idx[..i..], ..] += v[..i..], sequential in i
 * (crates/burn-ndarray/src/ops/base.rs:140-183) */
API void o_scatter_add_f32(float *t, const int64_t *idx, const float *v, size_t outer, size_t D,
                           size_t DI, size_t inner) {
  for (size_t o = 0; o < outer; ++o)
    for (size_t c = 0; c < inner; ++c)
      for (size_t i = 0; i < DI; ++i) {
        const size_t k = (o * DI + i) * inner + c;
        float *dst = &t[(o * D + (size_t)idx[k]) * inner + c];
        *dst = *dst + v[k];
      }
}

/* select (index_select) with 1-D indices of length NI */
API void o_select_f32(const float *t, const int64_t *idx, float *out, size_t outer, size_t D,
                      size_t NI, size_t inner) {
  for (size_t o = 0; o < outer; ++o)
    for (size_t i = 0; i < NI; ++i)
      memcpy(out + (o * NI + i) * inner, t + (o * D + (size_t)idx[i]) * inner, sizeof(float) * inner);
}

/* select_add (index_add): t[.., idx[i], ..] += v[.., i, ..], sequential in i */
API void o_select_add_f32(float *t, const int64_t *idx, const float *v, size_t outer, size_t D,
                          size_t NI, size_t inner) {
  for (size_t o = 0; o < outer; ++o)
    for (size_t i = 0; i < NI; ++i) {
      float *dst = t + (o * D + (size_t)idx[i]) * inner;
      const float *src = v + (o * NI + i) * inner;
      for (size_t c = 0; c < inner; ++c) dst[c] = dst[c] + src[c];
    }
}

/* ---- CPU baseline workload (bench.py cpu_baseline / --impl reference) ---- */
/*
 * The fused-chain benchmark of BASELINE.json configs[1], executed the way
 * burn-ndarray executes it: one full pass per primitive op, each producing a
 * new array (no fusion):
 *   t = a*b; t = t+c; u = t/sqrt2; u = erf(u); u = u+1; t = t*u; t = t/2;
 *   out = mask_fill(t, m, 0)
 * gelu primitives: crates/burn-backend/src/backend/ops/activation.rs:69-76.
 * tmp must hold 2*n floats.
 */
API void o_bench_chain_unfused(const float *a, const float *b, const float *c, const uint8_t *m,
                               float *out, float *tmp, size_t n) {
  float *t = tmp, *u = tmp + n;
  const float sqrt2 = (float)1.4142135623730951;
  o_binary_f32(O_MUL, a, b, t, n);
  o_binary_f32(O_ADD, t, c, t, n);
  o_binary_scalar_f32(O_DIV, t, sqrt2, u, n);
  o_unary_f32(U_ERF, u, u, n);
  o_binary_scalar_f32(O_ADD, u, 1.0f, u, n);
  o_binary_f32(O_MUL, t, u, t, n);
  o_binary_scalar_f32(O_DIV, t, 2.0f, t, n);
  o_mask_fill_f32(t, m, 0.0f, out, n);
}

/*
 * The same configs[1] step as ONE pass over the data — what a hand-fused CPU implementation would do, NOT what
 * burn-ndarray does (it has no fusion); reported beside the op-by-op figure as BASELINE.md §3 promises.  Per element:
 * y = mask_fill(gelu(a*b+c), m, 0) with the same roundings as the op-by-op chain (erf in f64, then f32), y stored,
 * and on the fly: row sum, row argmax (first max wins), column sums (accumulated into col_sum — the caller zeroes it
 * and, when sharding rows over threads, adds the per-thread partials), block total in f64.
 */
API void o_bench_step_fused(const float *a, const float *b, const float *c, const uint8_t *m, float *y,
                            float *row_sum, float *row_mean, int64_t *row_argmax, float *col_sum, double *total,
                            size_t rows, size_t cols) {
  const float sqrt2 = (float)1.4142135623730951;
  double tot = 0.0;
  for (size_t r = 0; r < rows; ++r) {
    const size_t base = r * cols;
    float acc = 0.0f, best = 0.0f;
    int64_t best_i = 0;
    for (size_t j = 0; j < cols; ++j) {
      const float t = a[base + j] * b[base + j] + c[base + j];
      float u = (float)erf((double)(t / sqrt2));
      u = u + 1.0f;
      float v = (t * u) / 2.0f;
      if (m[base + j]) v = 0.0f;
      y[base + j] = v;
      acc += v;
      col_sum[j] += v;
      if (j == 0 || v > best) { best = v; best_i = (int64_t)j; }
    }
    row_sum[r] = acc;
    row_mean[r] = acc / (float)cols;
    row_argmax[r] = best_i;
    tot += (double)acc;
  }
  *total = tot;
}
