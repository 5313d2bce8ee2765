// Gradient all-reduce over NCCL (NVLink 5 / NVSwitch).
//
// Replaces CubeBackend::all_reduce → ComputeClient::all_reduce and sync_collective
// (crates/burn-cubecl/src/ops/distributed.rs:17-55) behind
// DistributedOps::{all_reduce, sync_collective}
// (crates/burn-backend/src/backend/distributed/ops.rs:116-131).
//
// B200 design: one process per GPU; each process owns one ncclComm_t created
// from a unique id that the host bootstrap (torch.distributed / any store)
// distributes.  Collectives run in place on a dedicated high-priority stream,
// fenced by events against the producer (backward) stream so the all-reduce of
// early gradients overlaps the rest of backward; b200_collective_sync makes the
// consumer (optimizer) stream wait.  Small tensors are bucketed inside one
// ncclGroup (the reference issues one collective per parameter).
//
// libnccl is resolved at run time with dlopen so the core library loads on
// boxes without NCCL; a missing NCCL fails loudly at b200_comm_* time.
#include <dlfcn.h>

#include <cstring>

#include <mutex>
#include <vector>

#include "common.cuh"

namespace b200 {

// Minimal NCCL ABI (nccl.h 2.x) — declared locally so the build does not need
// the header at a matching version.
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5,
               ncclFloat16 = 6, ncclFloat32 = 7, ncclFloat64 = 8, ncclBfloat16 = 9 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3, ncclAvg = 4 } ncclRedOp_t;

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;
static std::mutex g_nccl_mu;

static int32_t load_nccl() {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.handle) return B200_OK;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  void *h = nullptr;
  for (const char *n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return fail(B200_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define B200_SYM(field, name)                                                     \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name));        \
  if (!g_nccl.field) return fail(B200_ERR_NCCL, "libnccl lacks symbol %s", name);
  B200_SYM(GetUniqueId, "ncclGetUniqueId")
  B200_SYM(CommInitRank, "ncclCommInitRank")
  B200_SYM(CommInitAll, "ncclCommInitAll")
  B200_SYM(CommDestroy, "ncclCommDestroy")
  B200_SYM(AllReduce, "ncclAllReduce")
  B200_SYM(GroupStart, "ncclGroupStart")
  B200_SYM(GroupEnd, "ncclGroupEnd")
  B200_SYM(GetErrorString, "ncclGetErrorString")
#undef B200_SYM
  g_nccl.handle = h;
  return B200_OK;
}

#define B200_NCCL(expr)                                                              \
  do {                                                                               \
    ncclResult_t _r = (expr);                                                        \
    if (_r != ncclSuccess)                                                           \
      return fail(B200_ERR_NCCL, "NCCL error %d (%s) in `%s`", (int)_r,              \
                  g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?", #expr);   \
  } while (0)

struct Comm {
  ncclComm_t comm = nullptr;
  cudaStream_t stream = nullptr;  // dedicated collective stream
  cudaEvent_t fence = nullptr;    // producer → collective stream
  cudaEvent_t done = nullptr;     // collective stream → consumer
  int rank = 0, world = 1;
  int device = -1;                // >= 0: made by b200_comm_init_all for that device (single-process multi-device)
};

static int32_t nccl_dtype(int32_t dt, ncclDataType_t *out) {
  switch (dt) {
    case B200_F32: *out = ncclFloat32; return B200_OK;
    case B200_F16: *out = ncclFloat16; return B200_OK;
    case B200_BF16: *out = ncclBfloat16; return B200_OK;
    case B200_I32: *out = ncclInt32; return B200_OK;
    case B200_I64: *out = ncclInt64; return B200_OK;
    default: return fail(B200_ERR_INVALID, "all_reduce: unsupported dtype %d", dt);
  }
}

}  // namespace b200

using namespace b200;

extern "C" int32_t b200_comm_unique_id(uint8_t id[B200_NCCL_UNIQUE_ID_BYTES]) {
  B200_REQUIRE(id, B200_ERR_INVALID, "id is null");
  int32_t st = load_nccl();
  if (st != B200_OK) return st;
  ncclUniqueId uid;
  B200_NCCL(g_nccl.GetUniqueId(&uid));
  memcpy(id, uid.internal, B200_NCCL_UNIQUE_ID_BYTES);
  return B200_OK;
}

extern "C" int32_t b200_comm_init(b200_comm *out, const uint8_t id[B200_NCCL_UNIQUE_ID_BYTES], int32_t rank,
                                  int32_t world_size) {
  B200_REQUIRE(out && id, B200_ERR_INVALID, "null argument");
  B200_REQUIRE(world_size >= 1 && rank >= 0 && rank < world_size, B200_ERR_INVALID, "bad rank %d / world %d", rank,
               world_size);
  int32_t st = load_nccl();
  if (st != B200_OK) return st;
  Comm *c = new Comm();
  c->rank = rank;
  c->world = world_size;
  ncclUniqueId uid;
  memcpy(uid.internal, id, B200_NCCL_UNIQUE_ID_BYTES);
  // a failure below must not leak the half-built communicator, its stream or its events
  auto build = [&]() -> int32_t {
    B200_NCCL(g_nccl.CommInitRank(&c->comm, world_size, uid, rank));
    int lo = 0, hi = 0;
    B200_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    B200_CUDA(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, coll_stream_priority(lo, hi)));
    B200_CUDA(cudaEventCreateWithFlags(&c->fence, cudaEventDisableTiming));
    B200_CUDA(cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming));
    return B200_OK;
  };
  st = build();
  if (st != B200_OK) {
    b200_comm_destroy((b200_comm)c);
    return st;
  }
  *out = (b200_comm)c;
  return B200_OK;
}

extern "C" int32_t b200_comm_destroy(b200_comm comm) {
  if (!comm) return B200_OK;
  Comm *c = (Comm *)comm;
  int prev = -1;
  if (c->device >= 0) {
    cudaGetDevice(&prev);
    cudaSetDevice(c->device);
  }
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->comm) g_nccl.CommDestroy(c->comm);
  if (c->fence) cudaEventDestroy(c->fence);
  if (c->done) cudaEventDestroy(c->done);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (prev >= 0) cudaSetDevice(prev);
  delete c;
  return B200_OK;
}

extern "C" int32_t b200_comm_init_all(b200_comm *out, const int32_t *devices, int32_t n) {
  B200_REQUIRE(out && devices && n >= 1 && n <= 64, B200_ERR_INVALID, "bad arguments to b200_comm_init_all");
  int32_t st = load_nccl();
  if (st != B200_OK) return st;
  std::vector<ncclComm_t> raw(n, nullptr);
  std::vector<int> devs(devices, devices + n);
  B200_NCCL(g_nccl.CommInitAll(raw.data(), n, devs.data()));
  int prev = 0;
  cudaGetDevice(&prev);
  std::vector<Comm *> cs(n, nullptr);
  auto build = [&]() -> int32_t {
    for (int i = 0; i < n; ++i) {
      Comm *c = new Comm();
      cs[i] = c;
      c->comm = raw[i];
      raw[i] = nullptr;
      c->rank = i;
      c->world = n;
      c->device = devs[i];
      B200_CUDA(cudaSetDevice(devs[i]));
      int lo = 0, hi = 0;
      B200_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
      B200_CUDA(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, coll_stream_priority(lo, hi)));
      B200_CUDA(cudaEventCreateWithFlags(&c->fence, cudaEventDisableTiming));
      B200_CUDA(cudaEventCreateWithFlags(&c->done, cudaEventDisableTiming));
    }
    return B200_OK;
  };
  st = build();
  cudaSetDevice(prev);
  if (st != B200_OK) {
    for (int i = 0; i < n; ++i) {
      if (cs[i]) b200_comm_destroy((b200_comm)cs[i]);
      else if (raw[i]) g_nccl.CommDestroy(raw[i]);
    }
    return st;
  }
  for (int i = 0; i < n; ++i) out[i] = (b200_comm)cs[i];
  return B200_OK;
}

extern "C" int32_t b200_all_reduce_group(const b200_comm *comms, void *const *ptrs, uint64_t count, int32_t n,
                                         int32_t dtype, int32_t op, void *const *producers) {
  B200_REQUIRE(comms && ptrs && n >= 1, B200_ERR_INVALID, "null argument");
  B200_REQUIRE(op == B200_REDUCE_SUM || op == B200_REDUCE_MEAN, B200_ERR_INVALID, "bad reduce op %d", op);
  ncclDataType_t dt;
  int32_t st = nccl_dtype(dtype, &dt);
  if (st != B200_OK) return st;
  if (count == 0) return B200_OK;
  int prev = 0;
  cudaGetDevice(&prev);
  const ncclRedOp_t rop = op == B200_REDUCE_MEAN ? ncclAvg : ncclSum;
  auto issue = [&]() -> int32_t {
    for (int i = 0; i < n; ++i) {
      Comm *c = (Comm *)comms[i];
      B200_REQUIRE(c && c->device >= 0, B200_ERR_INVALID, "communicator %d was not made by b200_comm_init_all", i);
      B200_CUDA(cudaSetDevice(c->device));
      B200_CUDA(cudaEventRecord(c->fence, producers ? (cudaStream_t)producers[i] : (cudaStream_t) nullptr));
      B200_CUDA(cudaStreamWaitEvent(c->stream, c->fence, 0));
    }
    // ONE group for all devices: a single thread issuing N blocking-order collectives outside a group deadlocks
    B200_NCCL(g_nccl.GroupStart());
    for (int i = 0; i < n; ++i) {
      Comm *c = (Comm *)comms[i];
      ncclResult_t r = g_nccl.AllReduce(ptrs[i], ptrs[i], (size_t)count, dt, rop, c->comm, c->stream);
      if (r != ncclSuccess) {
        g_nccl.GroupEnd();
        return fail(B200_ERR_NCCL, "NCCL error %d (%s) in ncclAllReduce for device %d", (int)r, g_nccl.GetErrorString(r), c->device);
      }
    }
    B200_NCCL(g_nccl.GroupEnd());
    return B200_OK;
  };
  st = issue();
  cudaSetDevice(prev);
  if (st == B200_OK) count_launch(n);
  return st;
}

extern "C" int32_t b200_comm_host_sync(b200_comm comm) {
  B200_REQUIRE(comm, B200_ERR_INVALID, "comm is null");
  Comm *c = (Comm *)comm;
  int prev = -1;
  if (c->device >= 0) {
    cudaGetDevice(&prev);
    cudaSetDevice(c->device);
  }
  cudaError_t e = cudaStreamSynchronize(c->stream);
  if (prev >= 0) cudaSetDevice(prev);
  B200_CUDA(e);
  return B200_OK;
}

extern "C" int32_t b200_all_reduce_multi(b200_comm comm, void *const *ptrs, const uint64_t *counts, int32_t n,
                                         int32_t dtype, int32_t op, b200_stream producer) {
  B200_REQUIRE(comm && ptrs && counts && n >= 0, B200_ERR_INVALID, "null argument");
  B200_REQUIRE(op == B200_REDUCE_SUM || op == B200_REDUCE_MEAN, B200_ERR_INVALID, "bad reduce op %d", op);
  Comm *c = (Comm *)comm;
  ncclDataType_t dt;
  int32_t st = nccl_dtype(dtype, &dt);
  if (st != B200_OK) return st;
  if (n == 0) return B200_OK;
  // order after everything the producer stream has queued so far
  B200_CUDA(cudaEventRecord(c->fence, resolve_stream(producer)));
  B200_CUDA(cudaStreamWaitEvent(c->stream, c->fence, 0));
  const ncclRedOp_t rop = op == B200_REDUCE_MEAN ? ncclAvg : ncclSum;
  B200_NCCL(g_nccl.GroupStart());
  for (int i = 0; i < n; ++i) {
    if (counts[i] == 0) continue;
    B200_NCCL(g_nccl.AllReduce(ptrs[i], ptrs[i], (size_t)counts[i], dt, rop, c->comm, c->stream));
  }
  B200_NCCL(g_nccl.GroupEnd());
  count_launch(1);
  return B200_OK;
}

extern "C" int32_t b200_all_reduce(b200_comm comm, void *ptr, uint64_t count, int32_t dtype, int32_t op,
                                   b200_stream producer) {
  void *ptrs[1] = {ptr};
  uint64_t counts[1] = {count};
  return b200_all_reduce_multi(comm, ptrs, counts, 1, dtype, op, producer);
}

extern "C" int32_t b200_collective_sync(b200_comm comm, b200_stream consumer) {
  B200_REQUIRE(comm, B200_ERR_INVALID, "comm is null");
  Comm *c = (Comm *)comm;
  B200_CUDA(cudaEventRecord(c->done, c->stream));
  B200_CUDA(cudaStreamWaitEvent(resolve_stream(consumer), c->done, 0));
  return B200_OK;
}

extern "C" int32_t b200_collective_mark(b200_comm comm, b200_event e) {
  B200_REQUIRE(comm && e, B200_ERR_INVALID, "null argument");
  Comm *c = (Comm *)comm;
  B200_CUDA(cudaEventRecord((cudaEvent_t)e, c->stream));
  return B200_OK;
}

extern "C" int32_t b200_stream_wait_event(b200_stream s, b200_event e) {
  B200_REQUIRE(e, B200_ERR_INVALID, "event is null");
  B200_CUDA(cudaStreamWaitEvent(resolve_stream(s), (cudaEvent_t)e, 0));
  return B200_OK;
}
