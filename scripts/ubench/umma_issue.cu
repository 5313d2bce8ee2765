// Microbenchmark (sm_100a): what does ONE tcgen05.mma of the attention kernels' shape cost to issue, and what does an
// mbarrier hand-over between two warps cost?  One CTA; results in cycles (clock64 of the issuing warp).
//   a) N back-to-back [128 x 64 x 8] tf32 SS instructions, fixed operands, one commit + wait at the end
//   b) the same with the descriptors advancing like a real K loop (8 k-steps, two tiles)
//   c) TS form (A from tensor memory)
//   d) commit -> waiting warp wakes -> arrives -> issuer wakes: round trip through two mbarriers, per hop
#include <cstdio>
#include <cuda_runtime.h>
#include "../../burn_b200/csrc/tcgen05.cuh"
using namespace b200::mm;

__device__ __forceinline__ void umma_e(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p, q;\nelect.sync _|q, 0xffffffff;\nsetp.ne.b32 p, %4, 0;\n"
               "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_ts_e(uint32_t tmem_d, uint32_t ta, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p, q;\nelect.sync _|q, 0xffffffff;\nsetp.ne.b32 p, %4, 0;\n"
               "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d), "r"(ta), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit_e(uint64_t *bar) {
  asm volatile("{\n.reg .pred q;\nelect.sync _|q, 0xffffffff;\n@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n}\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__global__ void __launch_bounds__(128, 1) bench(long long *out, int N, int NN) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 96 * 1024);
  uint32_t *slot = reinterpret_cast<uint32_t *>(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<float *>(smem)[i] = 0.0f;
  if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(bars + i, 1); fence_barrier_init(); }
  if (warp == 1) tmem_alloc(slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 32768), bmn = smem_u32(smem + 65536);
  if (warp == 1) {
    const uint32_t idesc = idesc_tf32(NN);
    uint32_t ph = 0;
    // a) fixed operands
    long long t0 = clock64();
    for (int i = 0; i < N; ++i) umma_e(tmem, make_desc(a0, 16, 1024), make_desc(b0, 16, 1024), idesc, 1u);
    long long t1 = clock64();
    commit_e(bars); mbar_wait(bars, ph); ph ^= 1;
    long long t2 = clock64();
    if (lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    // b) advancing descriptors, two products per round like S and dP
    t0 = clock64();
    for (int i = 0; i < N / 16; ++i) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma_e(tmem + (i & 1) * 64, make_desc(a0 + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
               make_desc(b0 + (i & 1) * 16384 + (k >> 2) * 8192 + (k & 3) * 32, 16, 1024), idesc, k ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma_e(tmem + 128 + (i & 1) * 64, make_desc(a0 + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
               make_desc(b0 + (i & 1) * 16384 + (k >> 2) * 8192 + (k & 3) * 32, 16, 1024), idesc, k ? 1u : 0u);
    }
    t1 = clock64();
    commit_e(bars); mbar_wait(bars, ph); ph ^= 1;
    t2 = clock64();
    if (lane == 0) { out[2] = t1 - t0; out[3] = t2 - t0; }
    // c) TS form, MN-major B
    const uint32_t idesc_mn = idesc_tf32(64) | (1u << 16);
    t0 = clock64();
    for (int i = 0; i < N / 8; ++i) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma_ts_e(tmem + 256, tmem + (i & 1) * 64 + k * 8, make_desc(bmn + (k >> 2) * 8192 + (k & 3) * 1024, 4096, 512, 1), idesc_mn, 1u);
    }
    t1 = clock64();
    commit_e(bars); mbar_wait(bars, ph); ph ^= 1;
    t2 = clock64();
    if (lane == 0) { out[4] = t1 - t0; out[5] = t2 - t0; }
    // d) one [128x64x64] product (8 instructions) + commit, then wait for the other warp's answer: round trips
    uint32_t ph1 = 0;
    t0 = clock64();
    for (int i = 0; i < 256; ++i) {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma_e(tmem, make_desc(a0 + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024), make_desc(b0 + (k >> 2) * 8192 + (k & 3) * 32, 16, 1024), idesc, k ? 1u : 0u);
      commit_e(bars + 2);
      mbar_wait(bars + 3, ph1); ph1 ^= 1;
    }
    t1 = clock64();
    if (lane == 0) out[6] = t1 - t0;
    // e) commit only (no MMA) round trips
    t0 = clock64();
    for (int i = 0; i < 256; ++i) {
      commit_e(bars + 2);
      mbar_wait(bars + 3, ph1); ph1 ^= 1;
    }
    t1 = clock64();
    if (lane == 0) out[7] = t1 - t0;
  } else if (warp == 2) {
    uint32_t ph = 0;
    for (int i = 0; i < 512; ++i) {
      mbar_wait(bars + 2, ph); ph ^= 1;
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 3);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long *d, h[8];
  cudaMalloc(&d, sizeof(h));
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024 + 1024);
  for (int nn : {64, 128, 256}) {
    const int N = 4096;
    for (int rep = 0; rep < 2; ++rep) {
      bench<<<1, 128, 100 * 1024 + 1024>>>(d, N, nn);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("N=%d (instruction N dim)  a) fixed operands: issue %.1f cyc/mma, issue+drain %.1f | b) K-loop descriptors: issue %.1f, +drain %.1f | "
           "c) TS form N=64: issue %.1f, +drain %.1f | d) 8 mma + commit + hand-over round trip %.0f cyc | e) commit + round trip %.0f cyc\n",
           nn, (double)h[0] / N, (double)h[1] / N, (double)h[2] / N, (double)h[3] / N, (double)h[4] / N, (double)h[5] / N, (double)h[6] / 256, (double)h[7] / 256);
  }
  return 0;
}
