// Microbenchmark (sm_100a): cost of a CTA's FIRST cp.async.bulk.tensor — issue time and landing time — for CTAs of the
// first and of later waves, with the tensor map (a) in the __grid_constant__ kernel parameter, (b) in global memory.
#include <cstdio>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include "../../burn_b200/csrc/tcgen05.cuh"
using namespace b200;
using namespace b200::mm;

struct Params { CUtensorMap map[4]; };

template <bool GLOBAL_MAP, bool PREFETCH>
__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ Params P, const CUtensorMap *gmaps, long long *out, int nmaps) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 64 * 1024);
  const CUtensorMap *maps = GLOBAL_MAP ? gmaps : P.map;
  if (threadIdx.x == 0) {
    if (PREFETCH) for (int i = 0; i < nmaps; ++i) tma_prefetch_desc(maps + i);
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    mbar_expect_tx(bar, 8192u * nmaps);
    long long ti[4];
    for (int i = 0; i < nmaps; ++i) {
      tma_load_5d(smem + i * 16384, maps + i, bar, 0, (blockIdx.x * 64) % 4096, 0, 0, 0);
      ti[i] = clock64() - t0;
    }
    mbar_wait(bar, 0);
    const long long t2 = clock64() - t0;
    out[blockIdx.x * 8 + 0] = ti[0];
    out[blockIdx.x * 8 + 1] = ti[nmaps - 1];
    out[blockIdx.x * 8 + 2] = t2;
  }
  // keep the CTA alive for a while so that waves are distinct
  if (threadIdx.x == 32) { const long long t = clock64(); while (clock64() - t < 20000) { } }
  __syncthreads();
}

int main() {
  const int rows = 4096, cols = 64;
  float *d; cudaMalloc(&d, (size_t)4 * rows * cols * 4);
  cudaMemset(d, 0, (size_t)4 * rows * cols * 4);
  Params P;
  for (int i = 0; i < 4; ++i) {
    Operand o; o.ptr = d + (size_t)i * rows * cols; o.es = 4; o.mn_major = false; o.s_mn = cols; o.s_k = 1;
    o.s_b[0] = o.s_b[1] = o.s_b[2] = 0; o.bsz[0] = o.bsz[1] = o.bsz[2] = 1;
    if (make_tmap(&P.map[i], o, rows, cols, 64) != 0) { printf("tmap failed: %s\n", b200_last_error()); return 1; }
  }
  CUtensorMap *gm; cudaMalloc(&gm, sizeof(P.map)); cudaMemcpy(gm, P.map, sizeof(P.map), cudaMemcpyHostToDevice);
  const int ctas = 148 * 6;
  long long *out; cudaMalloc(&out, ctas * 8 * sizeof(long long));
  std::vector<long long> h(ctas * 8);
  auto run = [&](const char *name, auto kern, int nmaps) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int rep = 0; rep < 2; ++rep) kern<<<ctas, 128, 200 * 1024>>>(P, gm, out, nmaps);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    cudaMemcpy(h.data(), out, h.size() * 8, cudaMemcpyDeviceToHost);
    auto med = [&](int lo, int hi, int f) { std::vector<long long> v; for (int i = lo; i < hi; ++i) v.push_back(h[i * 8 + f]); std::sort(v.begin(), v.end()); return v[v.size() / 2]; };
    printf("%-44s maps %d | wave 1: first issue %5lld, last issue %5lld, landed %5lld | waves 3-6: first issue %5lld, last issue %5lld, landed %5lld cycles\n",
           name, nmaps, med(0, 148, 0), med(0, 148, 1), med(0, 148, 2), med(296, ctas, 0), med(296, ctas, 1), med(296, ctas, 2));
  };
  for (int nm : {1, 4}) {
    run("param map, prefetch.tensormap", k<false, true>, nm);
    run("param map, no prefetch", k<false, false>, nm);
    run("global-memory map, prefetch.tensormap", k<true, true>, nm);
    run("global-memory map, no prefetch", k<true, false>, nm);
  }
  return 0;
}
