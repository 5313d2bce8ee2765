// Shared host/device helpers for libburn_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "burn_b200.h"

namespace b200 {

// ---- host-side error plumbing (runtime.cu) --------------------------------
int32_t fail(int32_t status, const char *fmt, ...);
int32_t fail_cuda(cudaError_t e, const char *what, const char *file, int line);
cudaStream_t resolve_stream(b200_stream s);
void count_launch(int n = 1);
int sm_count();
// Priority of the collective streams: highest by default (a bucket's sync starts the moment it is complete);
// B200_COLL_PRIORITY=low makes them background streams — compute kernels get free SM slots first.
static inline int coll_stream_priority(int lo, int hi) {
  const char *e = getenv("B200_COLL_PRIORITY");
  return (e && e[0] == 'l') ? lo : hi;
}
int max_smem_optin();
// sticky device-side error flag (host-mapped pinned int32): kernels store one of these codes, the next
// synchronising ABI call reports it as B200_ERR_SHAPE and clears it
enum : int32_t { kIdxErrGather = 1, kIdxErrSelect = 2, kIdxErrScatter = 3, kIdxErrSelectAdd = 4, kIdxErrTarget = 5, kPeerTimeout = 6 };
int32_t *index_error_flag();
int32_t check_index_error();
int32_t ensure_dyn_smem(const void *func, size_t bytes, bool max_carveout = false);

#define B200_CUDA(expr)                                                        \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess) return ::b200::fail_cuda(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define B200_REQUIRE(cond, status, ...)                                        \
  do {                                                                         \
    if (!(cond)) return ::b200::fail((status), __VA_ARGS__);                    \
  } while (0)

#define B200_LAUNCH_CHECK()                                                    \
  do {                                                                         \
    cudaError_t _e = cudaGetLastError();                                       \
    if (_e != cudaSuccess) return ::b200::fail_cuda(_e, "kernel launch", __FILE__, __LINE__); \
    ::b200::count_launch();                                                    \
  } while (0)

static inline int dtype_size(int32_t dt) {
  switch (dt) {
    case B200_F32: return 4;
    case B200_F16: return 2;
    case B200_BF16: return 2;
    case B200_I32: return 4;
    case B200_I64: return 8;
    case B200_BOOL: return 1;
    case B200_U8: return 1;
    default: return 0;
  }
}

static inline int64_t numel_of(const int64_t *shape, int rank) {
  int64_t n = 1;
  for (int i = 0; i < rank; ++i) n *= shape[i];
  return n;
}

static inline bool is_contiguous(const b200_tensor &t) {
  int64_t expect = 1;
  for (int d = t.rank - 1; d >= 0; --d) {
    if (t.shape[d] != 1 && t.strides[d] != expect) return false;
    expect *= t.shape[d];
  }
  return true;
}

// ---- 32-bit fast division by a runtime-constant divisor -------------------
// Valid for dividends n < 2^31 (callers pick the 64-bit path otherwise).
struct FastDiv {
  uint32_t d, magic, shift;
};

static inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d;
  if (d == 1) {
    f.magic = 0;
    f.shift = 0;
    return f;
  }
  uint32_t s = 0;
  while ((1ull << s) < d) ++s;
  uint64_t m = ((1ull << 32) * ((1ull << s) - d)) / d + 1;
  f.magic = (uint32_t)m;
  f.shift = s;
  return f;
}

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t fd_div(uint32_t n, const FastDiv &f) {
  // d == 1 → magic 0, shift 0 → (0 + n) >> 0 = n
  return (__umulhi(n, f.magic) + n) >> f.shift;
}
#endif

}  // namespace b200
