// Runtime plumbing of libburn_b200: device selection, streams, the stream-ordered
// caching allocator, host<->device copies and error reporting.
//
// Replaces the pieces of cubecl's CUDA runtime that CubeBackend leans on
// (crates/burn-cubecl/src/backend.rs:44-233 — name/seed/sync/memory_cleanup;
//  crates/burn-cubecl/src/tensor/base.rs:20-33 — handle ownership).
#include <atomic>
#include <cstdarg>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace b200 {

static thread_local char g_err[1024] = {0};
static std::atomic<uint64_t> g_launches{0};

struct DeviceState {
  bool ready = false;
  int sm_count = 0;
  int smem_optin = 0;
  cudaMemPool_t pool = nullptr;
  cudaStream_t default_stream = nullptr;
};
static DeviceState g_dev[16];
static std::mutex g_mu;

int32_t fail(int32_t status, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return status;
}

int32_t fail_cuda(cudaError_t e, const char *what, const char *file, int line) {
  snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s:%d in `%s`", (int)e,
           cudaGetErrorString(e), file, line, what);
  return B200_ERR_CUDA;
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

static int current_device() {
  int d = 0;
  cudaGetDevice(&d);
  return d;
}

static int32_t ensure_device(int d) {
  if (d < 0 || d >= 16) return fail(B200_ERR_INVALID, "device index %d out of range", d);
  std::lock_guard<std::mutex> lk(g_mu);
  DeviceState &st = g_dev[d];
  if (st.ready) return B200_OK;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(B200_ERR_NO_DEVICE, "no CUDA device available (%s) — burn-b200 has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (d >= n) return fail(B200_ERR_INVALID, "device %d requested but only %d present", d, n);
  B200_CUDA(cudaSetDevice(d));
  cudaDeviceProp prop;
  B200_CUDA(cudaGetDeviceProperties(&prop, d));
  if (prop.major != 10)
    return fail(B200_ERR_UNSUPPORTED,
                "device %d is sm_%d%d; libburn_b200 is built for sm_100a only", d, prop.major,
                prop.minor);
  st.sm_count = prop.multiProcessorCount;
  st.smem_optin = (int)prop.sharedMemPerBlockOptin;
  B200_CUDA(cudaDeviceGetDefaultMemPool(&st.pool, d));
  uint64_t keep = UINT64_MAX;  // caching behaviour: never trim until memory_cleanup
  B200_CUDA(cudaMemPoolSetAttribute(st.pool, cudaMemPoolAttrReleaseThreshold, &keep));
  B200_CUDA(cudaStreamCreateWithFlags(&st.default_stream, cudaStreamNonBlocking));
  st.ready = true;
  return B200_OK;
}

cudaStream_t resolve_stream(b200_stream s) {
  if (s) return (cudaStream_t)s;
  int d = current_device();
  if (!g_dev[d].ready) ensure_device(d);
  return g_dev[d].default_stream;
}

int sm_count() {
  int d = current_device();
  if (!g_dev[d].ready) ensure_device(d);
  return g_dev[d].sm_count > 0 ? g_dev[d].sm_count : 148;
}

int max_smem_optin() {
  int d = current_device();
  if (!g_dev[d].ready) ensure_device(d);
  return g_dev[d].smem_optin;
}

static int32_t *g_idx_err = nullptr;
int32_t *index_error_flag() {
  static std::once_flag once;
  std::call_once(once, [] {
    void *p = nullptr;
    if (cudaHostAlloc(&p, 64, cudaHostAllocPortable | cudaHostAllocMapped) == cudaSuccess) {
      g_idx_err = reinterpret_cast<int32_t *>(p);
      *g_idx_err = 0;
    }
  });
  return g_idx_err;   // UVA: the same pointer is valid on every device
}
int32_t check_index_error() {
  if (!g_idx_err) return B200_OK;
  const int32_t code = *reinterpret_cast<volatile int32_t *>(g_idx_err);
  if (code == 0) return B200_OK;
  *reinterpret_cast<volatile int32_t *>(g_idx_err) = 0;
  if (code == kPeerTimeout)
    return fail(B200_ERR_NCCL, "a peer-memory collective gave up waiting for another rank (10 s): ranks issued different "
                               "collectives, or a rank died; the buffers it touched are not valid");
  static const char *const what[] = {"?", "gather", "select", "scatter_add", "select_add", "softmax_cross_entropy targets"};
  return fail(B200_ERR_SHAPE, "an index was out of range in an earlier %s launch (the reference panics; the access was skipped)",
              what[code >= 1 && code <= 5 ? code : 0]);
}

// cudaFuncSetAttribute is a driver round trip (~1-2 us); a kernel's opt-in shared-memory limit only ever
// needs raising, so remember the largest value set per (function, device) and skip the call afterwards.
int32_t ensure_dyn_smem(const void *func, size_t bytes, bool max_carveout) {
  struct Entry { const void *f; int dev; size_t bytes; };
  static std::vector<Entry> table;
  static std::mutex mu;
  const int d = current_device();
  std::lock_guard<std::mutex> lk(mu);
  for (Entry &e : table)
    if (e.f == func && e.dev == d) {
      if (e.bytes >= bytes) return B200_OK;
      B200_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
      e.bytes = bytes;
      return B200_OK;
    }
  B200_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  if (max_carveout)
    B200_CUDA(cudaFuncSetAttribute(func, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  table.push_back({func, d, bytes});
  return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" {

int32_t b200_abi_version(void) { return B200_ABI_VERSION; }

const char *b200_last_error(void) { return g_err; }

int32_t b200_device_count(int32_t *count) {
  B200_REQUIRE(count, B200_ERR_INVALID, "count is null");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *count = 0;
    return fail(B200_ERR_NO_DEVICE, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
  }
  *count = n;
  return B200_OK;
}

int32_t b200_init(int32_t device) {
  int32_t st = ensure_device(device);
  if (st != B200_OK) return st;
  B200_CUDA(cudaSetDevice(device));
  return B200_OK;
}

int32_t b200_set_device(int32_t device) { return b200_init(device); }

int32_t b200_device_info(int32_t device, int32_t *sm, int32_t *major, int32_t *minor,
                         uint64_t *total_mem) {
  cudaDeviceProp prop;
  B200_CUDA(cudaGetDeviceProperties(&prop, device));
  if (sm) *sm = prop.multiProcessorCount;
  if (major) *major = prop.major;
  if (minor) *minor = prop.minor;
  if (total_mem) *total_mem = (uint64_t)prop.totalGlobalMem;
  return B200_OK;
}

int32_t b200_stream_create(b200_stream *out, int32_t high_priority) {
  B200_REQUIRE(out, B200_ERR_INVALID, "out is null");
  int lo = 0, hi = 0;
  B200_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  cudaStream_t s;
  B200_CUDA(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, high_priority ? hi : lo));
  *out = (b200_stream)s;
  return B200_OK;
}

int32_t b200_stream_destroy(b200_stream s) {
  if (s) B200_CUDA(cudaStreamDestroy((cudaStream_t)s));
  return B200_OK;
}

int32_t b200_stream_sync(b200_stream s) {
  B200_CUDA(cudaStreamSynchronize(resolve_stream(s)));
  return check_index_error();
}

int32_t b200_device_sync(void) {
  B200_CUDA(cudaDeviceSynchronize());
  return check_index_error();
}

int32_t b200_event_create(b200_event *out) {
  B200_REQUIRE(out, B200_ERR_INVALID, "out is null");
  cudaEvent_t e;
  B200_CUDA(cudaEventCreate(&e));
  *out = (b200_event)e;
  return B200_OK;
}

int32_t b200_event_destroy(b200_event e) {
  if (e) B200_CUDA(cudaEventDestroy((cudaEvent_t)e));
  return B200_OK;
}

int32_t b200_event_record(b200_event e, b200_stream s) {
  B200_CUDA(cudaEventRecord((cudaEvent_t)e, resolve_stream(s)));
  return B200_OK;
}

/* 1 = every launch recorded before `e` has finished, 0 = still running: the poll behind an async
 * float_into_data Future (crates/burn-backend/src/backend/ops/tensor.rs:109-111). */
int32_t b200_event_query(b200_event e, int32_t *done) {
  B200_REQUIRE(e && done, B200_ERR_INVALID, "null argument");
  cudaError_t r = cudaEventQuery((cudaEvent_t)e);
  if (r == cudaSuccess) { *done = 1; return check_index_error(); }
  if (r == cudaErrorNotReady) { *done = 0; return B200_OK; }
  return fail_cuda(r, "cudaEventQuery", __FILE__, __LINE__);
}

int32_t b200_event_elapsed_ms(b200_event start, b200_event stop, float *ms) {
  B200_REQUIRE(ms, B200_ERR_INVALID, "ms is null");
  B200_CUDA(cudaEventSynchronize((cudaEvent_t)stop));
  B200_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
  return B200_OK;
}

// ---------------------------------------------------------------- CUDA graphs
struct B200Graph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  uint64_t kernel_nodes = 0, total_nodes = 0;
};

int32_t b200_graph_begin(b200_stream s) {
  // relaxed: other host threads (and host-only driver calls such as tensor-map encoding) stay legal
  B200_CUDA(cudaStreamBeginCapture(resolve_stream(s), cudaStreamCaptureModeRelaxed));
  return B200_OK;
}

int32_t b200_graph_end(b200_stream s, b200_graph *out) {
  B200_REQUIRE(out, B200_ERR_INVALID, "out is null");
  B200Graph *g = new B200Graph();
  cudaError_t e = cudaStreamEndCapture(resolve_stream(s), &g->graph);
  if (e != cudaSuccess || !g->graph) {
    delete g;
    return fail(B200_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
  }
  size_t n = 0;
  B200_CUDA(cudaGraphGetNodes(g->graph, nullptr, &n));
  std::vector<cudaGraphNode_t> nodes(n);
  if (n) B200_CUDA(cudaGraphGetNodes(g->graph, nodes.data(), &n));
  g->total_nodes = n;
  for (size_t i = 0; i < n; ++i) {
    cudaGraphNodeType t;
    B200_CUDA(cudaGraphNodeGetType(nodes[i], &t));
    if (t == cudaGraphNodeTypeKernel) ++g->kernel_nodes;
  }
  // allocations still live at the end of one replay are released at the start of the next
  e = cudaGraphInstantiateWithFlags(&g->exec, g->graph, cudaGraphInstantiateFlagAutoFreeOnLaunch);
  if (e != cudaSuccess) {
    cudaGraphDestroy(g->graph);
    delete g;
    return fail(B200_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
  }
  *out = (b200_graph)g;
  return B200_OK;
}

int32_t b200_graph_launch(b200_graph gh, b200_stream s) {
  B200Graph *g = (B200Graph *)gh;
  B200_REQUIRE(g && g->exec, B200_ERR_INVALID, "graph is null");
  B200_CUDA(cudaGraphLaunch(g->exec, resolve_stream(s)));
  g_launches.fetch_add(g->kernel_nodes);
  return B200_OK;
}

int32_t b200_graph_destroy(b200_graph gh) {
  B200Graph *g = (B200Graph *)gh;
  if (!g) return B200_OK;
  if (g->exec) cudaGraphExecDestroy(g->exec);
  if (g->graph) cudaGraphDestroy(g->graph);
  delete g;
  return B200_OK;
}

int32_t b200_graph_node_count(b200_graph gh, uint64_t *kernel_nodes, uint64_t *total_nodes) {
  B200Graph *g = (B200Graph *)gh;
  B200_REQUIRE(g, B200_ERR_INVALID, "graph is null");
  if (kernel_nodes) *kernel_nodes = g->kernel_nodes;
  if (total_nodes) *total_nodes = g->total_nodes;
  return B200_OK;
}

// Handle refcounts (Handle::can_mut, crates/burn-ir/src/handle.rs:92-111): b200_alloc hands out a buffer with
// one owner, b200_retain adds one, b200_free drops one and releases the memory (stream-ordered) at zero.
static std::unordered_map<void *, uint32_t> g_refs;
static std::mutex g_refs_mu;

int32_t b200_alloc(void **out, uint64_t bytes, b200_stream s) {
  B200_REQUIRE(out, B200_ERR_INVALID, "out is null");
  if (bytes == 0) bytes = 16;  // zero-sized tensors still get a distinct handle
  B200_CUDA(cudaMallocAsync(out, (size_t)bytes, resolve_stream(s)));
  std::lock_guard<std::mutex> lk(g_refs_mu);
  g_refs[*out] = 1;
  return B200_OK;
}

int32_t b200_retain(void *ptr) {
  B200_REQUIRE(ptr, B200_ERR_INVALID, "ptr is null");
  std::lock_guard<std::mutex> lk(g_refs_mu);
  auto it = g_refs.find(ptr);
  B200_REQUIRE(it != g_refs.end(), B200_ERR_INVALID, "b200_retain: %p is not a live b200_alloc allocation", ptr);
  ++it->second;
  return B200_OK;
}

int32_t b200_refcount(const void *ptr, uint32_t *count) {
  B200_REQUIRE(ptr && count, B200_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(g_refs_mu);
  auto it = g_refs.find(const_cast<void *>(ptr));
  B200_REQUIRE(it != g_refs.end(), B200_ERR_INVALID, "b200_refcount: %p is not a live b200_alloc allocation", ptr);
  *count = it->second;
  return B200_OK;
}

int32_t b200_free(void *ptr, b200_stream s) {
  if (!ptr) return B200_OK;
  {
    std::lock_guard<std::mutex> lk(g_refs_mu);
    auto it = g_refs.find(ptr);
    if (it != g_refs.end()) {
      if (--it->second > 0) return B200_OK;   // other owners remain
      g_refs.erase(it);
    }
  }
  B200_CUDA(cudaFreeAsync(ptr, resolve_stream(s)));
  return B200_OK;
}

int32_t b200_memory_cleanup(void) {
  int d = current_device();
  B200_CUDA(cudaDeviceSynchronize());
  if (g_dev[d].ready) B200_CUDA(cudaMemPoolTrimTo(g_dev[d].pool, 0));
  return B200_OK;
}

int32_t b200_memset(void *ptr, int32_t byte, uint64_t bytes, b200_stream s) {
  B200_CUDA(cudaMemsetAsync(ptr, byte, (size_t)bytes, resolve_stream(s)));
  return B200_OK;
}

int32_t b200_host_alloc(void **out, uint64_t bytes) {
  B200_REQUIRE(out, B200_ERR_INVALID, "out is null");
  B200_CUDA(cudaMallocHost(out, (size_t)(bytes ? bytes : 16)));
  return B200_OK;
}

int32_t b200_host_free(void *ptr) {
  if (ptr) B200_CUDA(cudaFreeHost(ptr));
  return B200_OK;
}

int32_t b200_memcpy_h2d(void *dst, const void *src, uint64_t bytes, b200_stream s) {
  if (bytes == 0) return B200_OK;
  B200_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, resolve_stream(s)));
  return B200_OK;
}

int32_t b200_memcpy_d2h(void *dst, const void *src, uint64_t bytes, b200_stream s) {
  if (bytes == 0) return B200_OK;
  B200_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, resolve_stream(s)));
  return B200_OK;
}

int32_t b200_memcpy_d2d(void *dst, const void *src, uint64_t bytes, b200_stream s) {
  if (bytes == 0) return B200_OK;
  B200_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, resolve_stream(s)));
  return B200_OK;
}

uint64_t b200_launch_count(void) { return g_launches.load(); }
void b200_launch_count_reset(void) { g_launches.store(0); }

}  // extern "C"
