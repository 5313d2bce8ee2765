// Peer-memory collectives over NVLink 5 / NVSwitch: the gradient all-reduce of DistributedOps::all_reduce
// (crates/burn-backend/src/backend/distributed/ops.rs:116-131, crates/burn-cubecl/src/ops/distributed.rs:17-55) and —
// fused into the same kernel — the Adam update that consumes it (crates/burn-optim/src/optim/adam.rs:149-210).
//
// Every rank maps every other rank's region (cudaIpc handles between processes, cudaDeviceEnablePeerAccess inside one
// process).  One kernel per gradient bucket, identical on every rank, with rank r owning 1/N of the bucket:
//
//   arrive     block 0 stores the launch's epoch into every peer's flag word (st.release.sys); every block spins on
//              its OWN rank's flag words until all N ranks have arrived (ld.acquire.sys) — the gradients of all ranks
//              are complete and visible
//   reduce     for each 16-byte vector of the owned shard: N loads, one from every rank's gradient bucket over
//              NVLink (ld.global.cg — L2 only), summed in rank order 0..N-1 (the same order whoever owns the shard:
//              replicas stay bit-identical), scaled for Mean
//   update     MODE_ADAM: m, v (kept for the owned shard only — optimizer state is sharded N ways) and p are updated
//              with the reference's op sequence, the new parameter vector is stored into ALL N ranks' parameter
//              buckets (reduce-scatter → Adam on 1/N → all-gather in one pass; 7/8 of the optimizer work disappears)
//              MODE_REDUCE: the reduced vector is stored into all N gradient buckets (a two-shot all-reduce)
//   depart     the last block to finish stores the epoch into every peer's done word and waits for all N done words:
//              when the kernel ends, every shard of this rank's bucket has been written by its owner
//
// The kernel is built to run BESIDE the persistent tcgen05 GEMMs of the backward pass, not between them: 128 threads,
// <= 80 registers, no shared memory — 10 240 registers and one reserved KiB per CTA, which is what a 320-thread x 168-
// register, 225 KiB GEMM CTA leaves free on an SM.  NCCL's ring kernels cannot co-reside and time-slice with the GEMMs
// (round-1 finding: +2.3 ms per step at N=2 with ~2 ms of wire time).  Epochs live in device memory and advance in the
// kernel itself, so a captured CUDA graph replays without host patching.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace b200 {
namespace peer {

constexpr int kMaxRanks = B200_PEER_MAX_RANKS;
constexpr uint64_t kFlagBytes = 256u << 10;  // head of every region
constexpr int kSlotWords = 32;               // [0..7] arrive, [8..15] done, [16] epoch, [17] finished-block counter
constexpr int kMaxSlots = (int)(kFlagBytes / (kSlotWords * 4));
constexpr int kThreads = 128;

enum { MODE_REDUCE = 0, MODE_ADAM = 1 };

struct Args {
  float *data[kMaxRanks];      // data base (after the flag area) of every rank's region, as mapped in THIS process
  uint32_t *flags[kMaxRanks];  // flag area of every rank's region
  uint64_t g_off, p_off;       // element offsets of the gradient / parameter bucket inside the data area
  float *m, *v;                // local moments, full-bucket indexing (only the owned shard is touched)
  const float *coef;           // device [2]: cf, eps_t
  uint64_t n;                  // elements in the bucket (multiple of 4)
  int rank, world, slot;
  int mean, pow2;
  float inv_world, world_f;
  float lr, b1, b2, omb1, omb2;
  int32_t *err;                // the library's sticky error word (host-mapped)
};

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t now_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// Bounded: a rank that never arrives (mismatched launch sequences, a dead process) must not wedge the GPU.  After
// 10 s the waiter records kPeerTimeout in the library's sticky error word (reported by the next synchronising ABI
// call) and carries on; whatever it then computes is garbage and is flagged as such.
__device__ __forceinline__ void wait_for(const uint32_t *p, uint32_t epoch, int32_t *err) {
  if ((int32_t)(ld_acquire_sys(p) - epoch) >= 0) return;
  const uint64_t t0 = now_ns();
  while ((int32_t)(ld_acquire_sys(p) - epoch) < 0) {
    __nanosleep(100);
    if (now_ns() - t0 > 10000000000ull) {
      *reinterpret_cast<volatile int32_t *>(err) = kPeerTimeout;
      return;
    }
  }
}

__device__ __forceinline__ void adam1(float &p, float &m, float &v, float g, const Args &A, float cf, float eps_t) {
  m = __fadd_rn(__fmul_rn(m, A.b1), __fmul_rn(g, A.omb1));
  v = __fadd_rn(__fmul_rn(v, A.b2), __fmul_rn(__fmul_rn(g, g), A.omb2));
  const float u = __fdiv_rn(__fmul_rn(m, cf), __fadd_rn(__fsqrt_rn(v), eps_t));
  p = __fsub_rn(p, __fmul_rn(u, A.lr));
}

template <int W, int MODE>
__global__ void __launch_bounds__(kThreads, 6) peer_bucket_kernel(const Args A) {
  uint32_t *mine = A.flags[A.rank] + (size_t)A.slot * kSlotWords;
  const uint32_t epoch = *reinterpret_cast<volatile uint32_t *>(mine + 16) + 1u;
  const int tid = threadIdx.x;

  // ---- arrive
  if (blockIdx.x == 0 && tid < W) {
    __threadfence_system();
    st_release_sys(A.flags[tid] + (size_t)A.slot * kSlotWords + A.rank, epoch);
  }
  if (tid < W) wait_for(mine + tid, epoch, A.err);
  __syncthreads();

  // ---- reduce (+ update) the owned shard
  const uint64_t n4 = A.n >> 2;
  const uint64_t per = (n4 + W - 1) / W;
  const uint64_t lo = (uint64_t)A.rank * per;
  const uint64_t hi = lo + per < n4 ? lo + per : n4;
  float cf = 0.f, eps_t = 0.f;
  if (MODE == MODE_ADAM) {
    cf = __ldg(A.coef);
    eps_t = __ldg(A.coef + 1);
  }
  // U vectors per thread per trip so that U x W = 8 sixteen-byte loads are in flight per thread whatever the world size:
  // an NVLink round trip is ~2-3 us, and 148 CTAs x 128 threads x 128 B = 2.4 MB in flight is what keeps ~0.8 TB/s busy
  constexpr int U = W <= 2 ? 4 : (W <= 4 ? 2 : 1);
  const uint64_t tpg = (uint64_t)gridDim.x * kThreads;
  for (uint64_t i0 = lo + (uint64_t)blockIdx.x * kThreads + tid; i0 < hi; i0 += tpg * U) {
    float4 part[U][W];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t i = i0 + u * tpg;
      if (i < hi) {
#pragma unroll
        for (int k = 0; k < W; ++k) part[u][k] = __ldcg(reinterpret_cast<const float4 *>(A.data[k] + A.g_off) + i);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint64_t i = i0 + u * tpg;
      if (i >= hi) break;
      float4 g = part[u][0];
#pragma unroll
      for (int k = 1; k < W; ++k) {
        g.x = __fadd_rn(g.x, part[u][k].x);
        g.y = __fadd_rn(g.y, part[u][k].y);
        g.z = __fadd_rn(g.z, part[u][k].z);
        g.w = __fadd_rn(g.w, part[u][k].w);
      }
      if (A.mean) {
        if (A.pow2) {
          g.x = __fmul_rn(g.x, A.inv_world); g.y = __fmul_rn(g.y, A.inv_world);
          g.z = __fmul_rn(g.z, A.inv_world); g.w = __fmul_rn(g.w, A.inv_world);
        } else {
          g.x = __fdiv_rn(g.x, A.world_f); g.y = __fdiv_rn(g.y, A.world_f);
          g.z = __fdiv_rn(g.z, A.world_f); g.w = __fdiv_rn(g.w, A.world_f);
        }
      }
      if (MODE == MODE_ADAM) {
        float4 p = __ldcg(reinterpret_cast<const float4 *>(A.data[A.rank] + A.p_off) + i);
        float4 m = __ldcs(reinterpret_cast<const float4 *>(A.m) + i), v = __ldcs(reinterpret_cast<const float4 *>(A.v) + i);
        adam1(p.x, m.x, v.x, g.x, A, cf, eps_t);
        adam1(p.y, m.y, v.y, g.y, A, cf, eps_t);
        adam1(p.z, m.z, v.z, g.z, A, cf, eps_t);
        adam1(p.w, m.w, v.w, g.w, A, cf, eps_t);
        __stcs(reinterpret_cast<float4 *>(A.m) + i, m);
        __stcs(reinterpret_cast<float4 *>(A.v) + i, v);
#pragma unroll
        for (int k = 0; k < W; ++k) __stcg(reinterpret_cast<float4 *>(A.data[k] + A.p_off) + i, p);
      } else {
#pragma unroll
        for (int k = 0; k < W; ++k) __stcg(reinterpret_cast<float4 *>(A.data[k] + A.g_off) + i, g);
      }
    }
  }

  // ---- depart: the last block tells every peer this rank's shard is in place, then waits for theirs
  __threadfence_system();
  __syncthreads();
  int last = 0;
  if (tid == 0) last = atomicAdd(mine + 17, 1u) == gridDim.x - 1;
  last = __syncthreads_or(last);
  if (!last) return;
  if (tid == 0) mine[17] = 0;
  __threadfence_system();
  if (tid < W) {
    st_release_sys(A.flags[tid] + (size_t)A.slot * kSlotWords + 8 + A.rank, epoch);
    wait_for(mine + 8 + tid, epoch, A.err);
  }
  __syncthreads();
  if (tid == 0) {
    *reinterpret_cast<volatile uint32_t *>(mine + 16) = epoch;
    __threadfence();
  }
}

struct Group {
  int rank = 0, world = 1, device = 0;
  bool ipc = false;                    // peers were opened with cudaIpcOpenMemHandle
  bool owns_local = false;             // create_local: the region itself belongs to the group
  void *base[kMaxRanks] = {};          // region base of every rank as mapped here
  uint64_t bytes = 0;
  cudaStream_t stream = nullptr;       // dedicated high-priority collective stream
  cudaEvent_t fence = nullptr, done = nullptr;
};

template <int MODE>
static int32_t launch(Group *g, const Args &A, cudaStream_t stream) {
  // four CTAs per SM: beside a persistent GEMM only one of them fits at a time (the others follow as it finishes);
  // beside lighter kernels, or alone at the end of backward, all four run (256 MiB at N=2: 1.00 / 0.55 / 0.46 ms with 1 / 2 / 4)
  const uint64_t per = ((A.n >> 2) + A.world - 1) / A.world;
  const int u = A.world <= 2 ? 4 : (A.world <= 4 ? 2 : 1);
  static const int ctas_per_sm = [] { const char *e = std::getenv("B200_PEER_CTAS_PER_SM"); const int v = e ? atoi(e) : 4; return v >= 1 && v <= 16 ? v : 4; }();
  const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((per + kThreads * u - 1) / (kThreads * u), (uint64_t)ctas_per_sm * sm_count()));
  switch (A.world) {
#define B200_PEER_CASE(W) \
  case W: peer_bucket_kernel<W, MODE><<<grid, kThreads, 0, stream>>>(A); break;
    B200_PEER_CASE(1) B200_PEER_CASE(2) B200_PEER_CASE(3) B200_PEER_CASE(4)
    B200_PEER_CASE(5) B200_PEER_CASE(6) B200_PEER_CASE(7) B200_PEER_CASE(8)
#undef B200_PEER_CASE
    default: return fail(B200_ERR_INVALID, "peer group of %d ranks (limit %d)", A.world, kMaxRanks);
  }
  (void)g;
  B200_LAUNCH_CHECK();
  return B200_OK;
}

static int32_t fill_common(Group *g, Args &A, uint64_t count, int32_t slot, int32_t op) {
  B200_REQUIRE(slot >= 0 && slot < kMaxSlots, B200_ERR_INVALID, "flag slot %d out of range (0..%d)", slot, kMaxSlots - 1);
  B200_REQUIRE(count % 4 == 0, B200_ERR_INVALID, "peer collectives work on multiples of 4 elements (got %llu)", (unsigned long long)count);
  B200_REQUIRE(op == B200_REDUCE_SUM || op == B200_REDUCE_MEAN, B200_ERR_INVALID, "bad reduce op %d", op);
  memset(&A, 0, sizeof(A));
  for (int k = 0; k < g->world; ++k) {
    A.flags[k] = reinterpret_cast<uint32_t *>(g->base[k]);
    A.data[k] = reinterpret_cast<float *>(reinterpret_cast<char *>(g->base[k]) + kFlagBytes);
  }
  A.n = count;
  A.rank = g->rank;
  A.world = g->world;
  A.slot = slot;
  A.mean = op == B200_REDUCE_MEAN;
  A.pow2 = (g->world & (g->world - 1)) == 0;
  A.inv_world = 1.0f / (float)g->world;
  A.world_f = (float)g->world;
  A.err = index_error_flag();
  return B200_OK;
}

static int32_t make_streams(Group *g) {
  int lo = 0, hi = 0;
  B200_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  B200_CUDA(cudaStreamCreateWithPriority(&g->stream, cudaStreamNonBlocking, coll_stream_priority(lo, hi)));
  B200_CUDA(cudaEventCreateWithFlags(&g->fence, cudaEventDisableTiming));
  B200_CUDA(cudaEventCreateWithFlags(&g->done, cudaEventDisableTiming));
  return B200_OK;
}

// launches on g's device whatever the calling thread's current device is (single-process multi-device groups)
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

}  // namespace peer
}  // namespace b200

using namespace b200;
using peer::Group;

extern "C" int32_t b200_peer_alloc(void **out, uint64_t bytes) {
  B200_REQUIRE(out && bytes > 0, B200_ERR_INVALID, "bad arguments to b200_peer_alloc");
  // cudaMalloc, not the stream-ordered pool: legacy IPC handles cannot export pool memory
  B200_CUDA(cudaMalloc(out, bytes));
  B200_CUDA(cudaMemset(*out, 0, bytes));
  B200_CUDA(cudaDeviceSynchronize());
  return B200_OK;
}

extern "C" int32_t b200_peer_free(void *ptr) {
  if (ptr) B200_CUDA(cudaFree(ptr));
  return B200_OK;
}

extern "C" int32_t b200_peer_export(void *ptr, uint8_t handle[B200_PEER_HANDLE_BYTES]) {
  B200_REQUIRE(ptr && handle, B200_ERR_INVALID, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == B200_PEER_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t h;
  B200_CUDA(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle, &h, sizeof(h));
  return B200_OK;
}

extern "C" uint64_t b200_peer_flag_bytes(void) { return peer::kFlagBytes; }

extern "C" int32_t b200_peer_group_create(b200_peer_group *out, int32_t rank, int32_t world, void *local, uint64_t bytes,
                                          const uint8_t *handles) {
  B200_REQUIRE(out && local, B200_ERR_INVALID, "null argument");
  B200_REQUIRE(world >= 1 && world <= peer::kMaxRanks && rank >= 0 && rank < world, B200_ERR_INVALID, "bad rank %d / world %d", rank, world);
  B200_REQUIRE(bytes > peer::kFlagBytes, B200_ERR_INVALID, "a peer region starts with %llu flag bytes", (unsigned long long)peer::kFlagBytes);
  B200_REQUIRE(world == 1 || handles, B200_ERR_INVALID, "handles is null");
  Group *g = new Group();
  g->rank = rank;
  g->world = world;
  g->bytes = bytes;
  g->ipc = true;
  cudaGetDevice(&g->device);
  auto build = [&]() -> int32_t {
    for (int k = 0; k < world; ++k) {
      if (k == rank) {
        g->base[k] = local;
        continue;
      }
      cudaIpcMemHandle_t h;
      memcpy(&h, handles + (size_t)k * B200_PEER_HANDLE_BYTES, sizeof(h));
      B200_CUDA(cudaIpcOpenMemHandle(&g->base[k], h, cudaIpcMemLazyEnablePeerAccess));
    }
    return peer::make_streams(g);
  };
  int32_t st = build();
  if (st != B200_OK) {
    b200_peer_group_destroy((b200_peer_group)g);
    return st;
  }
  *out = (b200_peer_group)g;
  return B200_OK;
}

extern "C" int32_t b200_peer_group_create_local(b200_peer_group *out, const int32_t *devices, int32_t n, uint64_t bytes) {
  B200_REQUIRE(out && devices, B200_ERR_INVALID, "null argument");
  B200_REQUIRE(n >= 1 && n <= peer::kMaxRanks, B200_ERR_INVALID, "bad device count %d", n);
  B200_REQUIRE(bytes > peer::kFlagBytes, B200_ERR_INVALID, "a peer region starts with %llu flag bytes", (unsigned long long)peer::kFlagBytes);
  int prev = 0;
  cudaGetDevice(&prev);
  std::vector<Group *> gs(n, nullptr);
  std::vector<void *> regions(n, nullptr);
  auto build = [&]() -> int32_t {
    for (int i = 0; i < n; ++i) {
      B200_CUDA(cudaSetDevice(devices[i]));
      B200_CUDA(cudaMalloc(&regions[i], bytes));
      B200_CUDA(cudaMemset(regions[i], 0, bytes));
      for (int j = 0; j < n; ++j) {
        if (i == j) continue;
        int can = 0;
        B200_CUDA(cudaDeviceCanAccessPeer(&can, devices[i], devices[j]));
        B200_REQUIRE(can, B200_ERR_UNSUPPORTED, "device %d cannot access device %d", devices[i], devices[j]);
        cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        else if (e != cudaSuccess) return fail_cuda(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
      }
      B200_CUDA(cudaDeviceSynchronize());
    }
    for (int i = 0; i < n; ++i) {
      B200_CUDA(cudaSetDevice(devices[i]));
      Group *g = new Group();
      gs[i] = g;
      g->rank = i;
      g->world = n;
      g->device = devices[i];
      g->bytes = bytes;
      g->owns_local = true;
      for (int k = 0; k < n; ++k) g->base[k] = regions[k];
      int32_t st = peer::make_streams(g);
      if (st != B200_OK) return st;
    }
    return B200_OK;
  };
  int32_t st = build();
  cudaSetDevice(prev);
  if (st != B200_OK) {
    for (int i = 0; i < n; ++i) {
      if (gs[i]) b200_peer_group_destroy((b200_peer_group)gs[i]);
      else if (regions[i]) cudaFree(regions[i]);
    }
    return st;
  }
  for (int i = 0; i < n; ++i) out[i] = (b200_peer_group)gs[i];
  return B200_OK;
}

extern "C" void *b200_peer_data(b200_peer_group group) {
  Group *g = (Group *)group;
  return g ? reinterpret_cast<char *>(g->base[g->rank]) + peer::kFlagBytes : nullptr;
}

extern "C" int32_t b200_peer_group_destroy(b200_peer_group group) {
  if (!group) return B200_OK;
  Group *g = (Group *)group;
  peer::DeviceGuard guard(g->device);
  if (g->stream) cudaStreamSynchronize(g->stream);
  if (g->ipc)
    for (int k = 0; k < g->world; ++k)
      if (k != g->rank && g->base[k]) cudaIpcCloseMemHandle(g->base[k]);
  if (g->owns_local && g->base[g->rank]) cudaFree(g->base[g->rank]);
  if (g->fence) cudaEventDestroy(g->fence);
  if (g->done) cudaEventDestroy(g->done);
  if (g->stream) cudaStreamDestroy(g->stream);
  delete g;
  return B200_OK;
}

static int32_t enter(Group *g, b200_stream producer, bool own_default_stream) {
  // order the collective stream after everything the producer stream has queued so far
  cudaStream_t ps = own_default_stream ? resolve_stream(producer) : (cudaStream_t)producer;
  B200_CUDA(cudaEventRecord(g->fence, ps));
  B200_CUDA(cudaStreamWaitEvent(g->stream, g->fence, 0));
  return B200_OK;
}

extern "C" int32_t b200_launch_peer_all_reduce(b200_peer_group group, uint64_t offset, uint64_t count, int32_t op,
                                               int32_t slot, b200_stream producer) {
  B200_REQUIRE(group, B200_ERR_INVALID, "group is null");
  Group *g = (Group *)group;
  peer::Args A;
  int32_t st = peer::fill_common(g, A, count, slot, op);
  if (st != B200_OK) return st;
  B200_REQUIRE(offset % 4 == 0 && (offset + count) * 4 + peer::kFlagBytes <= g->bytes, B200_ERR_INVALID,
               "bucket [%llu, +%llu) does not fit the region", (unsigned long long)offset, (unsigned long long)count);
  if (count == 0) return B200_OK;
  A.g_off = offset;
  peer::DeviceGuard guard(g->device);
  // a group made by create_local may live on a device that is not this library's current one: its producer stream
  // is then a raw cudaStream_t of that device (nullptr = that device's legacy default stream)
  if ((st = enter(g, producer, !g->owns_local)) != B200_OK) return st;
  return peer::launch<peer::MODE_REDUCE>(g, A, g->stream);
}

extern "C" int32_t b200_launch_peer_adam(b200_peer_group group, uint64_t grad_offset, uint64_t param_offset,
                                         void *moment1, void *moment2, const void *coef, uint64_t count, double lr,
                                         double beta1, double beta2, int32_t slot, b200_stream producer) {
  B200_REQUIRE(group && moment1 && moment2 && coef, B200_ERR_INVALID, "null argument");
  Group *g = (Group *)group;
  peer::Args A;
  int32_t st = peer::fill_common(g, A, count, slot, B200_REDUCE_MEAN);
  if (st != B200_OK) return st;
  for (uint64_t off : {grad_offset, param_offset})
    B200_REQUIRE(off % 4 == 0 && (off + count) * 4 + peer::kFlagBytes <= g->bytes, B200_ERR_INVALID,
                 "bucket [%llu, +%llu) does not fit the region", (unsigned long long)off, (unsigned long long)count);
  B200_REQUIRE((((uintptr_t)moment1 | (uintptr_t)moment2) & 15) == 0, B200_ERR_INVALID, "moments must be 16-byte aligned");
  if (count == 0) return B200_OK;
  A.g_off = grad_offset;
  A.p_off = param_offset;
  A.m = reinterpret_cast<float *>(moment1);
  A.v = reinterpret_cast<float *>(moment2);
  A.coef = reinterpret_cast<const float *>(coef);
  A.lr = (float)lr;
  A.b1 = (float)beta1;
  A.b2 = (float)beta2;
  A.omb1 = 1.0f - (float)beta1;
  A.omb2 = 1.0f - (float)beta2;
  peer::DeviceGuard guard(g->device);
  if ((st = enter(g, producer, !g->owns_local)) != B200_OK) return st;
  return peer::launch<peer::MODE_ADAM>(g, A, g->stream);
}

extern "C" int32_t b200_peer_sync(b200_peer_group group, b200_stream consumer) {
  B200_REQUIRE(group, B200_ERR_INVALID, "group is null");
  Group *g = (Group *)group;
  peer::DeviceGuard guard(g->device);
  B200_CUDA(cudaEventRecord(g->done, g->stream));
  B200_CUDA(cudaStreamWaitEvent(g->owns_local ? (cudaStream_t)consumer : resolve_stream(consumer), g->done, 0));
  return B200_OK;
}

extern "C" int32_t b200_peer_mark(b200_peer_group group, b200_event e) {
  B200_REQUIRE(group && e, B200_ERR_INVALID, "null argument");
  Group *g = (Group *)group;
  B200_CUDA(cudaEventRecord((cudaEvent_t)e, g->stream));
  return B200_OK;
}

extern "C" int32_t b200_peer_host_sync(b200_peer_group group) {
  B200_REQUIRE(group, B200_ERR_INVALID, "group is null");
  Group *g = (Group *)group;
  peer::DeviceGuard guard(g->device);
  B200_CUDA(cudaStreamSynchronize(g->stream));
  return B200_OK;
}
