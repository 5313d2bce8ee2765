// NVRTC specialisation of hot elementwise tapes ("consider NVRTC specialisation as a
// second mode").  The reference JIT-compiles every fused trace (CubeCL → CUDA C++ → NVRTC,
// crates/burn-cubecl-fusion/src/engine/codegen/kernel.rs); here the op-tape interpreter stays the
// general path and large, linear launches get a kernel generated from the compiled tape: straight-
// line code over U=2 vectors of 4 elements per thread, operands in registers (no slot file, no
// dispatch), 128-bit streaming loads/stores.  Each op is `eval_op<OPC>` from tape_eval.cuh — the
// interpreter's arithmetic with the opcode folded at compile time — so results are bit-identical.
// Kernels are cached by generated source; libnvrtc is dlopen()ed, and when it is missing or a
// compile fails the launch simply stays on the interpreter (still a GPU path).
// B200_TAPE_JIT=0 disables, B200_TAPE_JIT_MIN_VEC overrides the size threshold.
#include <dlfcn.h>

#include <cstdlib>
#include <mutex>
#include <string>
#include <unordered_map>

#include "tape_host.cuh"
#include "tape_eval.cuh"
#include "jit_embed.inc"

namespace b200 {
namespace jit {

typedef struct _nvrtcProgram *nvrtcProgram;
struct Nvrtc {
  void *h = nullptr;
  int (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
  int (*CompileProgram)(nvrtcProgram, int, const char *const *) = nullptr;
  int (*GetCUBINSize)(nvrtcProgram, size_t *) = nullptr;
  int (*GetCUBIN)(nvrtcProgram, char *) = nullptr;
  int (*GetProgramLogSize)(nvrtcProgram, size_t *) = nullptr;
  int (*GetProgramLog)(nvrtcProgram, char *) = nullptr;
  int (*DestroyProgram)(nvrtcProgram *) = nullptr;
  bool ok = false, tried = false;
};
static Nvrtc g_rtc;
static std::mutex g_mu;
// Specialised kernels are shape-specialised, so a dynamic-shape workload keeps producing new ones: the cache is a
// bounded LRU (B200_JIT_CACHE_MAX entries, default 512).  Evicting unloads the cubin; kernels of it may still be in
// flight, so the device is drained first — never inside a stream capture, where eviction is simply postponed.
struct CacheEntry {
  cudaKernel_t kern = nullptr;   // nullptr = compile failed, do not retry
  cudaLibrary_t lib = nullptr;
  uint64_t tick = 0;
};
static std::unordered_map<std::string, CacheEntry> g_cache;
static uint64_t g_tick = 0, g_evictions = 0;
static size_t cache_max() {
  static const size_t v = [] {
    const char *e = std::getenv("B200_JIT_CACHE_MAX");
    const long n = e ? atol(e) : 512;
    return (size_t)(n >= 1 ? n : 512);
  }();
  return v;
}

static bool load_nvrtc() {
  if (g_rtc.tried) return g_rtc.ok;
  g_rtc.tried = true;
  for (const char *name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"}) {
    g_rtc.h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
    if (g_rtc.h) break;
  }
  if (!g_rtc.h) {
    fprintf(stderr, "[burn_b200] libnvrtc not found: elementwise tapes stay on the interpreter\n");
    return false;
  }
#define B200_RTC_SYM(field, sym)                                            \
  *(void **)(&g_rtc.field) = dlsym(g_rtc.h, sym);                           \
  if (!g_rtc.field) { fprintf(stderr, "[burn_b200] libnvrtc lacks %s\n", sym); return false; }
  B200_RTC_SYM(CreateProgram, "nvrtcCreateProgram")
  B200_RTC_SYM(CompileProgram, "nvrtcCompileProgram")
  B200_RTC_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
  B200_RTC_SYM(GetCUBIN, "nvrtcGetCUBIN")
  B200_RTC_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
  B200_RTC_SYM(GetProgramLog, "nvrtcGetProgramLog")
  B200_RTC_SYM(DestroyProgram, "nvrtcDestroyProgram")
#undef B200_RTC_SYM
  g_rtc.ok = true;
  return true;
}

// B200_TAPE_JIT_STRICT=1 (set by the test-suite): a tape that qualifies for specialisation but fails to
// compile or load is an error instead of a silent fall back to the interpreter.
static bool strict() {
  static const bool v = std::getenv("B200_TAPE_JIT_STRICT") != nullptr;
  return v;
}

static bool dtype_in_ok(int32_t dt) {
  return dt == B200_F32 || dt == B200_I32 || dt == B200_BOOL || dt == B200_U8 || dt == B200_BF16 || dt == B200_F16;
}
static bool dtype_out_ok(int32_t dt) { return dtype_in_ok(dt); }

// Generates the kernel source for a compiled tape over linear operands.
struct Tuning {
  int u, block, ctas_per_sm;
};
static const Tuning &tuning() {
  static const Tuning t = [] {
    auto env = [](const char *name, int dflt) { const char *e = std::getenv(name); return e ? atoi(e) : dflt; };
    Tuning x{env("B200_JIT_U", 2), env("B200_JIT_BLOCK", 256), env("B200_JIT_CTAS_PER_SM", 8)};
    if (x.u < 1 || x.u > 8) x.u = 2;
    if (x.block != 128 && x.block != 256 && x.block != 512) x.block = 256;
    if (x.ctas_per_sm < 1 || x.ctas_per_sm > 32) x.ctas_per_sm = 8;
    return x;
  }();
  return t;
}

// Generates the kernel source for a compiled tape.  rank3 == false: every operand is contiguous over the
// (collapsed, 1-D) iteration space or one broadcast element.  rank3 == true: collapsed rank <= 3 with
// per-operand strides (broadcast rows / columns, strided views); the shape and strides are baked into
// the source as literals so the index arithmetic strength-reduces (the kernel is shape-specialised,
// like every kernel the reference JIT-compiles).
static std::string generate(const CompiledTape &ct, const TapeParams &p, bool rank3) {
  std::string s;
  const std::string BLK = std::to_string(tuning().block) + "u";
  auto N = [](long long x) { return std::to_string(x); };
  s += "#include \"burn_b200.h\"\n#include \"tape_eval.cuh\"\nusing namespace b200;\n";
  s += "extern \"C\" __global__ void __launch_bounds__(" + N(tuning().block) + ") b200_jit_kernel(const JitParams P) {\n";
  s += "  constexpr int U = " + N(tuning().u) + ";\n  const uint32_t stride = gridDim.x * " + BLK + ";\n";
  if (rank3) s += "  constexpr uint32_t SH1 = " + N(p.shape[1]) + "u, SH2V = " + N(p.shape[2] / 4) + "u;\n";
  for (size_t k = 0; k < ct.scalars.size(); ++k) s += "  const uint32_t S" + N(k) + " = P.scalars[" + N(k) + "];\n";
  auto row_varying = [&](const OperandDesc &d) { return rank3 && (d.s3[0] != 0 || d.s3[1] != 0); };
  for (int i = 0; i < ct.n_in; ++i)
    if (p.in[i].mode == kModeBcast && !row_varying(p.in[i]))
      s += "  const uint32_t B" + N(i) + " = jit_ld1<" + N(p.in[i].dtype) + ">(P.in[" + N(i) + "]);\n";
  s += "  for (uint32_t v0 = blockIdx.x * " + BLK + " + threadIdx.x; v0 < P.n_vec; v0 += U * stride) {\n";
  for (int i = 0; i < ct.n_in; ++i) {
    if (p.in[i].mode == kModeVec) s += "    uint32_t I" + N(i) + "[U][4];\n";
    else if (row_varying(p.in[i])) s += "    uint32_t R" + N(i) + "[U];\n";
  }
  // vector offset of operand d at the current coordinates
  auto voff = [&](const OperandDesc &d) -> std::string {
    if (!rank3) return "v";
    return "(c0 * " + N(d.s3[0] / 4) + "u + c1 * " + N(d.s3[1] / 4) + "u + c2v)";
  };
  const std::string coords = rank3 ? "        const uint32_t c2v = v % SH2V, t_ = v / SH2V, c1 = t_ % SH1, c0 = t_ / SH1;\n" : "";
  s += "#pragma unroll\n    for (int u = 0; u < U; ++u) {\n      const uint32_t v = v0 + u * stride;\n      if (v < P.n_vec) {\n" + coords;
  for (int i = 0; i < ct.n_in; ++i) {
    if (p.in[i].mode == kModeVec)
      s += "        jit_ld4<" + N(p.in[i].dtype) + ">(P.in[" + N(i) + "], " + voff(p.in[i]) + ", I" + N(i) + "[u]);\n";
    else if (row_varying(p.in[i]))
      s += "        R" + N(i) + "[u] = jit_ld1_at<" + N(p.in[i].dtype) + ">(P.in[" + N(i) + "], c0 * " + N(p.in[i].s3[0]) + "u + c1 * " +
           N(p.in[i].s3[1]) + "u);\n";
  }
  s += "      }\n    }\n";
  s += "#pragma unroll\n    for (int u = 0; u < U; ++u) {\n      const uint32_t v = v0 + u * stride;\n      if (v < P.n_vec) {\n" + coords;
  for (int o = 0; o < ct.n_out; ++o) s += "        uint32_t O" + N(o) + "[4];\n";
  s += "#pragma unroll\n        for (int j = 0; j < 4; ++j) {\n          uint32_t a = 0u;\n";
  for (int t = 0; t < ct.n_tmp; ++t) s += "          uint32_t T" + N(t) + " = 0u;\n";
  auto arg = [&](const SymArg &x, bool acc_if_none) -> std::string {
    switch (x.kind) {
      case 1:
        if (p.in[x.idx].mode == kModeVec) return "I" + N(x.idx) + "[u][j]";
        return row_varying(p.in[x.idx]) ? "R" + N(x.idx) + "[u]" : "B" + N(x.idx);
      case 2: return "T" + N(x.idx);
      case 3: return "S" + N(x.idx);
      default: return acc_if_none ? "a" : "0u";
    }
  };
  for (const SymOp &o : ct.ops) {
    s += "          a = eval_op<" + N(o.op) + ">(a, " + arg(o.b, o.op == kOpGelu) + ", " + arg(o.c, false) + ");\n";
    if (o.dst_tmp >= 0) s += "          T" + N(o.dst_tmp) + " = a;\n";
    if (o.dst_out >= 0) s += "          O" + N(o.dst_out) + "[j] = a;\n";
  }
  s += "        }\n";
  for (int o = 0; o < ct.n_out; ++o)
    s += "        jit_st4<" + N(p.out[o].dtype) + ">(P.out[" + N(o) + "], " + voff(p.out[o]) + ", O" + N(o) + ");\n";
  s += "      }\n    }\n  }\n}\n";
  return s;
}

static cudaKernel_t compile(const std::string &src, const char *kernel_name = "b200_jit_kernel", bool load = true,
                            size_t *cubin_bytes = nullptr, cudaLibrary_t *lib_out = nullptr) {
  nvrtcProgram prog = nullptr;
  const char *hdr_src[] = {kJitSrc_burn_b200_h, kJitSrc_tape_eval_cuh, kJitSrc_tape_math_cuh, kJitSrc_erf_table_inc, kJitSrc_tanh_table_inc,
                           kJitSrc_stdint_h, kJitSrc_stdint_h};
  const char *hdr_name[] = {"burn_b200.h", "tape_eval.cuh", "tape_math.cuh", "erf_table.inc", "tanh_table.inc", "stdint.h", "stddef.h"};
  if (g_rtc.CreateProgram(&prog, src.c_str(), "b200_jit.cu", 7, hdr_src, hdr_name) != 0) return nullptr;
  // -default-device: burn_b200.h's host prototypes are only declarations here
  const char *opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo", "-default-device", "-diag-suppress=177"};
  const int rc = g_rtc.CompileProgram(prog, 5, opts);
  cudaKernel_t kern = nullptr;
  if (rc != 0) {
    size_t n = 0;
    g_rtc.GetProgramLogSize(prog, &n);
    std::string log(n + 1, '\0');
    if (n) g_rtc.GetProgramLog(prog, &log[0]);
    fprintf(stderr, "[burn_b200] NVRTC failed (%d); this tape stays on the interpreter:\n%.2000s\n", rc, log.c_str());
  } else {
    size_t n = 0;
    if (g_rtc.GetCUBINSize(prog, &n) == 0 && n) {
      if (cubin_bytes) *cubin_bytes = n;
      std::string cubin(n, '\0');
      if (load && g_rtc.GetCUBIN(prog, &cubin[0]) == 0) {
        cudaLibrary_t lib = nullptr;
        if (cudaLibraryLoadData(&lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0) == cudaSuccess) {
          if (cudaLibraryGetKernel(&kern, lib, kernel_name) != cudaSuccess) kern = nullptr;
          if (kern && lib_out) *lib_out = lib;
          else if (!kern) cudaLibraryUnload(lib);
        }
        if (!kern) {
          fprintf(stderr, "[burn_b200] loading a specialised kernel failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        }
      }
    }
  }
  g_rtc.DestroyProgram(&prog);
  return kern;
}

// g_mu held.  Looks the source up, compiling on a miss; keeps the cache within its bound.
static cudaKernel_t cached(const std::string &key, const std::string &src, const char *kernel_name, cudaStream_t stream) {
  auto it = g_cache.find(key);
  if (it == g_cache.end()) {
    if (g_cache.size() >= cache_max()) {
      cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
      if (cudaStreamIsCapturing(stream, &cs) != cudaSuccess) cudaGetLastError();
      if (cs == cudaStreamCaptureStatusNone) {
        auto victim = g_cache.begin();
        for (auto e = g_cache.begin(); e != g_cache.end(); ++e)
          if (e->second.tick < victim->second.tick) victim = e;
        if (victim->second.lib) {
          cudaDeviceSynchronize();
          cudaLibraryUnload(victim->second.lib);
        }
        g_cache.erase(victim);
        ++g_evictions;
      }
    }
    CacheEntry e;
    e.kern = compile(src, kernel_name, true, nullptr, &e.lib);
    it = g_cache.emplace(key, e).first;
  }
  it->second.tick = ++g_tick;
  return it->second.kern;
}

// ---- fuse-on-read reductions --------------------------------------------------------------
// The read tape evaluated per vector, folded locally and combined per CTA; one partial (or, when a
// row / column is not split, the final value) per CTA and lane.  Split partials are finished by the
// plain fast reduction, so the combine order is fixed.
struct RedParams {       // mirrored in the generated source
  float *out;            // final output or partial buffer
  uint32_t outer, R, inner4;     // R in elements for columns; rows use r4 = R/4 vectors
  uint32_t r4, splits, per_split;
  float div;             // mean divisor applied when the kernel writes final values (0 = none)
};

static std::string gen_reduce(const CompiledTape &ct, const TapeParams &p, int32_t kind, bool col) {
  std::string s;
  s += "#include \"burn_b200.h\"\n#include \"tape_eval.cuh\"\nusing namespace b200;\n";
  s += "struct RedParams { float *out; uint32_t outer, R, inner4; uint32_t r4, splits, per_split; float div; };\n";
  const bool is_sum = kind == B200_RED_SUM || kind == B200_RED_MEAN, is_max = kind == B200_RED_MAX;
  // NVRTC has no <math.h>: spell the infinities as bit patterns
  s += std::string("#define IDENT ") + (is_sum ? "0.0f" : (is_max ? "__int_as_float(0xff800000)" : "__int_as_float(0x7f800000)")) + "\n";
  if (is_sum) s += "__device__ __forceinline__ float comb(float a, float b) { return __fadd_rn(a, b); }\n";
  else s += std::string("__device__ __forceinline__ float comb(float a, float b) { return (a != a) ? a : ((b != b) ? b : ") +
            (is_max ? "fmaxf" : "fminf") + "(a, b)); }\n";
  // LOAD_EVAL(g, val): the read tape on the 4 elements of vector g
  s += "#define LOAD_EVAL(g, val) { \\\n";
  for (int i = 0; i < ct.n_in; ++i)
    if (p.in[i].mode == kModeVec)
      s += "  uint32_t I" + std::to_string(i) + "[4]; jit_ld4<" + std::to_string(p.in[i].dtype) + ">(P.in[" + std::to_string(i) + "], g, I" + std::to_string(i) + "); \\\n";
  s += "  _Pragma(\"unroll\") for (int j = 0; j < 4; ++j) { uint32_t a = 0u; \\\n";
  for (int t = 0; t < ct.n_tmp; ++t) s += "    uint32_t T" + std::to_string(t) + " = 0u; \\\n";
  auto arg = [&](const SymArg &x, bool acc_if_none) -> std::string {
    switch (x.kind) {
      case 1: return p.in[x.idx].mode == kModeBcast ? "B" + std::to_string(x.idx) : "I" + std::to_string(x.idx) + "[j]";
      case 2: return "T" + std::to_string(x.idx);
      case 3: return "S" + std::to_string(x.idx);
      default: return acc_if_none ? "a" : "0u";
    }
  };
  for (const SymOp &o : ct.ops) {
    s += "    a = eval_op<" + std::to_string(o.op) + ">(a, " + arg(o.b, o.op == kOpGelu) + ", " + arg(o.c, false) + "); \\\n";
    if (o.dst_tmp >= 0) s += "    T" + std::to_string(o.dst_tmp) + " = a; \\\n";
  }
  s += "    val[j] = f_of(a); } }\n";
  auto prelude = [&]() {
    std::string q;
    for (size_t k = 0; k < ct.scalars.size(); ++k) q += "  const uint32_t S" + std::to_string(k) + " = P.scalars[" + std::to_string(k) + "];\n";
    for (int i = 0; i < ct.n_in; ++i)
      if (p.in[i].mode == kModeBcast)
        q += "  const uint32_t B" + std::to_string(i) + " = jit_ld1<" + std::to_string(p.in[i].dtype) + ">(P.in[" + std::to_string(i) + "]);\n";
    return q;
  };
  if (!col) {
    // TPR threads per (row, split): 32 (a warp, shuffle only) or 256 (a CTA)
    s += "template <int TPR> __device__ __forceinline__ void rows(const JitParams &P, const RedParams &Q) {\n" + prelude();
    s += R"(  __shared__ float scratch[8];
  const uint32_t lane_in = threadIdx.x % TPR, grp = threadIdx.x / TPR, groups = 256 / TPR;
  const uint32_t n_work = Q.outer * Q.splits;
  for (uint32_t w0 = blockIdx.x * groups; w0 < n_work; w0 += gridDim.x * groups) {
    const uint32_t w = w0 + grp;
    float acc = IDENT;
    if (w < n_work) {
      const uint32_t row = w / Q.splits, split = w - row * Q.splits;
      const uint32_t begin = split * Q.per_split, end = min(Q.r4, begin + Q.per_split);
      for (uint32_t i0 = begin + lane_in; i0 < end; i0 += 2 * TPR) {
        float v0[4], v1[4];
        const uint32_t i1 = i0 + TPR;
        LOAD_EVAL(row * Q.r4 + i0, v0)
        if (i1 < end) { LOAD_EVAL(row * Q.r4 + i1, v1) } else { v1[0] = v1[1] = v1[2] = v1[3] = IDENT; }
        acc = comb(acc, comb(comb(v0[0], v0[1]), comb(v0[2], v0[3])));
        acc = comb(acc, comb(comb(v1[0], v1[1]), comb(v1[2], v1[3])));
      }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) acc = comb(acc, __shfl_xor_sync(0xffffffffu, acc, m));
    if (TPR == 256) {
      __syncthreads();
      if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = acc;
      __syncthreads();
      acc = (threadIdx.x & 31) < 8 ? scratch[threadIdx.x & 31] : IDENT;
#pragma unroll
      for (int m = 4; m >= 1; m >>= 1) acc = comb(acc, __shfl_xor_sync(0xffffffffu, acc, m));
    }
    if (lane_in == 0 && w < n_work) Q.out[w] = Q.div != 0.0f ? __fdiv_rn(acc, Q.div) : acc;
  }
}
extern "C" __global__ void __launch_bounds__(256) b200_jit_rows_warp(const JitParams P, const RedParams Q) { rows<32>(P, Q); }
extern "C" __global__ void __launch_bounds__(256) b200_jit_rows_cta(const JitParams P, const RedParams Q) { rows<256>(P, Q); }
)";
  } else {
    // 8 CTAs per SM (<= 32 registers): the fuse-on-read math is issue-bound, so resident warps are what it needs
    s += "extern \"C\" __global__ void __launch_bounds__(256, 8) b200_jit_cols(const JitParams P, const RedParams Q) {\n" + prelude();
    s += R"(  // block (32, 8): 32 column vectors x 8 row groups; grid (outer * column tiles, splits)
  __shared__ float part[8][32][4];
  const uint32_t tiles = (Q.inner4 + 31) / 32;
  const uint32_t o = blockIdx.x / tiles, c4 = (blockIdx.x - o * tiles) * 32 + threadIdx.x;
  const uint32_t r_begin = blockIdx.y * Q.per_split, r_end = min(Q.R, r_begin + Q.per_split);
  float acc[4] = {IDENT, IDENT, IDENT, IDENT};
  if (c4 < Q.inner4) {
    for (uint32_t r = r_begin + threadIdx.y; r < r_end; r += 16) {
      float v0[4], v1[4];
      const uint32_t r1 = r + 8;
      LOAD_EVAL((o * Q.R + r) * Q.inner4 + c4, v0)
      if (r1 < r_end) { LOAD_EVAL((o * Q.R + r1) * Q.inner4 + c4, v1) } else { v1[0] = v1[1] = v1[2] = v1[3] = IDENT; }
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[j] = comb(acc[j], comb(v0[j], v1[j]));
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) part[threadIdx.y][threadIdx.x][j] = acc[j];
  __syncthreads();
  if (threadIdx.y == 0 && c4 < Q.inner4) {
    float4 r;
    float *rr = reinterpret_cast<float *>(&r);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float b = part[0][threadIdx.x][j];
      for (int y = 1; y < 8; ++y) b = comb(b, part[y][threadIdx.x][j]);
      rr[j] = Q.div != 0.0f ? __fdiv_rn(b, Q.div) : b;
    }
    // partials laid out [outer][splits][inner]
    reinterpret_cast<float4 *>(Q.out)[((size_t)o * gridDim.y + blockIdx.y) * Q.inner4 + c4] = r;
  }
}
)";
  }
  return s;
}

}  // namespace jit

int32_t fast_reduce_f32(int32_t kind, int64_t outer, int64_t R, int64_t inner, const float *in, float *out,
                        float mean_div, cudaStream_t stream);

// Fuse-on-read reduce over linear operands into a contiguous f32 output.  1 = launched, 0 = not applicable.
int32_t jit_try_reduce(const CompiledTape &ct, const TapeParams &p, int rank_mode_, int32_t kind, int64_t outer,
                       int64_t R, int64_t inner, float *out, cudaStream_t stream) {
  static const bool enabled = [] { const char *e = std::getenv("B200_TAPE_JIT"); return !(e && e[0] == '0'); }();
  static const uint32_t min_vec = [] {
    const char *e = std::getenv("B200_TAPE_JIT_MIN_VEC");
    return e ? (uint32_t)strtoul(e, nullptr, 10) : (1u << 18);
  }();
  const bool col = inner > 1;
  if (!enabled || rank_mode_ != kRankLinear || p.n_vec < min_vec || R < 1) return 0;
  if (col ? (inner % 4 != 0) : (R % 4 != 0)) return 0;
  if (outer * R * inner >= (1ll << 32) || ((uintptr_t)out & 15)) return 0;
  for (int i = 0; i < ct.n_in; ++i) {
    const OperandDesc &d = p.in[i];
    if (!jit::dtype_in_ok(d.dtype)) return 0;
    if (!(d.mode == kModeBcast || (d.mode == kModeVec && d.s3[2] == 1))) return 0;
  }
  const int sms = sm_count();
  jit::RedParams Q;
  memset(&Q, 0, sizeof(Q));
  Q.outer = (uint32_t)outer;
  Q.R = (uint32_t)R;
  const char *kname;
  dim3 grid, block(256);
  uint32_t splits = 1;
  if (!col) {
    Q.r4 = (uint32_t)(R / 4);
    const bool warp_rows = Q.r4 <= 1024 && outer >= (int64_t)sms * 8;
    if (!warp_rows && outer < (int64_t)sms * 4 && Q.r4 > 4096)
      splits = (uint32_t)std::min<int64_t>((sms * 4 + outer - 1) / outer, Q.r4 / 2048);
    splits = std::max(1u, splits);
    Q.per_split = ((Q.r4 + splits - 1) / splits + 511) / 512 * 512;
    splits = (Q.r4 + Q.per_split - 1) / Q.per_split;
    kname = warp_rows ? "b200_jit_rows_warp" : "b200_jit_rows_cta";
    const uint64_t work = (uint64_t)outer * splits, per_cta = warp_rows ? 8 : 1;
    grid = dim3((unsigned)std::max<uint64_t>(1, std::min<uint64_t>((work + per_cta - 1) / per_cta, (uint64_t)sms * 8)));
  } else {
    Q.inner4 = (uint32_t)(inner / 4);
    const uint32_t tiles = (uint32_t)outer * ((Q.inner4 + 31) / 32);
    // >= 4 waves of 8 resident CTAs per SM: with a single partial wave (1024 CTAs on 148 x 8 slots) the SMs that drew
    // 6 CTAs idle while the others finish their 7th — ncu: 52 % warps active, 139 us against the row mapping's 108 us
    while (splits < 64 && tiles * splits < (uint32_t)sms * 32u && R / (splits * 2) >= 64) splits *= 2;
    Q.per_split = (uint32_t)((R + splits - 1) / splits);
    splits = (uint32_t)((R + Q.per_split - 1) / Q.per_split);
    kname = "b200_jit_cols";
    grid = dim3(tiles, splits);
    block = dim3(32, 8);
  }
  Q.splits = splits;
  const bool mean = kind == B200_RED_MEAN;
  Q.div = (splits == 1 && mean) ? (float)R : 0.0f;

  cudaKernel_t kern = nullptr;
  {
    std::lock_guard<std::mutex> lock(jit::g_mu);
    if (!jit::load_nvrtc()) return 0;
    const std::string src = jit::gen_reduce(ct, p, kind, col);
    const std::string key = src + "//" + kname;
    kern = jit::cached(key, src, kname, stream);
  }
  if (!kern) return jit::strict() ? fail(B200_ERR_CUDA, "NVRTC specialisation failed (see stderr) and B200_TAPE_JIT_STRICT is set") : 0;
  float *partials = nullptr;
  if (splits > 1) B200_CUDA(cudaMallocAsync((void **)&partials, (size_t)outer * splits * (col ? inner : 1) * sizeof(float), stream));
  Q.out = splits > 1 ? partials : out;
  JitParams P;
  memset(&P, 0, sizeof(P));
  for (int i = 0; i < ct.n_in; ++i) P.in[i] = p.in[i].ptr;
  for (size_t k = 0; k < ct.scalars.size(); ++k) P.scalars[k] = ct.scalars[k];
  P.n_vec = p.n_vec;
  void *args[] = {&P, &Q};
  B200_CUDA(cudaLaunchKernel((const void *)kern, grid, block, args, 0, stream));
  count_launch(1);
  int32_t st = 1;
  if (splits > 1) {
    // finish: plain reduction over the split axis ([outer, splits, inner]); mean divides by the full R
    const int32_t fs = fast_reduce_f32(kind, outer, splits, col ? inner : 1, partials, out, mean ? (float)R : 0.0f, stream);
    cudaFreeAsync(partials, stream);
    if (fs <= 0) st = fs < 0 ? fs : fail(B200_ERR_UNSUPPORTED, "no fast path to finish %u reduce partials", splits);
  }
  return st;
}

// Returns 1 when the launch ran on a specialised kernel, 0 when the caller should use the interpreter.
int32_t jit_try_elemwise(const CompiledTape &ct, const TapeParams &p, int vec, int rank_mode_, cudaStream_t stream) {
  static const bool enabled = [] { const char *e = std::getenv("B200_TAPE_JIT"); return !(e && e[0] == '0'); }();
  static const uint32_t min_vec = [] {
    const char *e = std::getenv("B200_TAPE_JIT_MIN_VEC");
    return e ? (uint32_t)strtoul(e, nullptr, 10) : (1u << 18);
  }();
  if (!enabled || vec != 4 || (rank_mode_ != kRankLinear && rank_mode_ != kRank3) || p.n_vec < min_vec) return 0;
  const bool rank3 = rank_mode_ == kRank3;
  auto strides_ok = [&](const OperandDesc &d, bool is_vec) {
    if (d.s3[0] < 0 || d.s3[1] < 0) return false;
    return !is_vec || !rank3 || (d.s3[0] % 4 == 0 && d.s3[1] % 4 == 0);
  };
  for (int i = 0; i < ct.n_in; ++i) {
    const OperandDesc &d = p.in[i];
    if (!jit::dtype_in_ok(d.dtype)) return 0;
    if (!(d.mode == kModeBcast || (d.mode == kModeVec && d.s3[2] == 1))) return 0;
    if (!strides_ok(d, d.mode == kModeVec)) return 0;
  }
  for (int o = 0; o < ct.n_out; ++o)
    if (!jit::dtype_out_ok(p.out[o].dtype) || p.out[o].mode != kModeVec || p.out[o].s3[2] != 1 || !strides_ok(p.out[o], true))
      return 0;

  cudaKernel_t kern = nullptr;
  {
    std::lock_guard<std::mutex> lock(jit::g_mu);
    if (!jit::load_nvrtc()) return 0;
    const std::string src = jit::generate(ct, p, rank3);
    kern = jit::cached(src, src, "b200_jit_kernel", stream);
  }
  if (!kern) return jit::strict() ? fail(B200_ERR_CUDA, "NVRTC specialisation failed (see stderr) and B200_TAPE_JIT_STRICT is set") : 0;
  JitParams P;
  memset(&P, 0, sizeof(P));
  for (int i = 0; i < ct.n_in; ++i) P.in[i] = p.in[i].ptr;
  for (int o = 0; o < ct.n_out; ++o) P.out[o] = p.out[o].ptr;
  for (size_t k = 0; k < ct.scalars.size(); ++k) P.scalars[k] = ct.scalars[k];
  P.n_vec = p.n_vec;
  const uint32_t per_block = (uint32_t)(jit::tuning().block * jit::tuning().u);
  const unsigned grid = (unsigned)std::max<uint32_t>(1u, std::min<uint32_t>((p.n_vec + per_block - 1) / per_block,
                                                                           (uint32_t)(sm_count() * jit::tuning().ctas_per_sm)));
  void *args[] = {&P};
  B200_CUDA(cudaLaunchKernel((const void *)kern, dim3(grid), dim3(jit::tuning().block), args, 0, stream));
  count_launch(1);
  return 1;
}

}  // namespace b200

// Generates and NVRTC-compiles (for sm_100a, without loading — no device needed) the specialised kernels of
// a reference tape: the bench chain as an elementwise kernel (linear and rank-3 forms) and as a fuse-on-read
// row / column reduction.  The "does the JIT path build" check of the CPU test-suite.
extern "C" int32_t b200_jit_selftest(uint64_t *cubin_bytes_total) {
  using namespace b200;
  b200_tape_op ops[8];
  memset(ops, 0, sizeof(ops));
  auto set = [&](int i, int op, uint8_t a, uint8_t b, uint8_t c, uint8_t dt, uint8_t dout) {
    ops[i].op = (uint8_t)op; ops[i].a = a; ops[i].b = b; ops[i].c = c; ops[i].dst_temp = dt; ops[i].dst_out = dout;
  };
  const uint8_t NONE = B200_DST_NONE;
  set(0, B200_OP_MUL_F, B200_ARG_INPUT(0), B200_ARG_INPUT(1), 0, NONE, NONE);
  set(1, B200_OP_ADD_F, B200_ARG_ACC, B200_ARG_INPUT(2), 0, 0, NONE);
  set(2, B200_OP_DIV_F, B200_ARG_TEMP(0), B200_ARG_SCALAR(0), 0, NONE, NONE);
  set(3, B200_OP_ERF_F, B200_ARG_ACC, 0, 0, NONE, NONE);
  set(4, B200_OP_ADD_F, B200_ARG_ACC, B200_ARG_SCALAR(1), 0, NONE, NONE);
  set(5, B200_OP_MUL_F, B200_ARG_TEMP(0), B200_ARG_ACC, 0, NONE, NONE);
  set(6, B200_OP_DIV_F, B200_ARG_ACC, B200_ARG_SCALAR(2), 0, NONE, NONE);
  set(7, B200_OP_SELECT, B200_ARG_ACC, B200_ARG_SCALAR(3), B200_ARG_INPUT(3), NONE, 0);
  const float sc[4] = {1.41421356237f, 1.0f, 2.0f, 0.0f};
  uint32_t scalars[4];
  memcpy(scalars, sc, sizeof(sc));
  b200_tape tape = {ops, 8, scalars, 4};
  CompiledTape ct;
  int32_t st = compile_tape(&tape, 4, 1, ct);
  if (st != B200_OK) return st;
  TapeParams p;
  memset(&p, 0, sizeof(p));
  for (int i = 0; i < 4; ++i) {
    p.in[i].mode = kModeVec;
    p.in[i].dtype = i == 3 ? B200_BOOL : B200_F32;
    p.in[i].s3[0] = 0; p.in[i].s3[1] = 4096; p.in[i].s3[2] = 1;
  }
  p.in[2].mode = kModeBcast;            // a per-row operand in the rank-3 form
  p.in[2].s3[1] = 1; p.in[2].s3[2] = 0;
  p.out[0].mode = kModeVec; p.out[0].dtype = B200_F32; p.out[0].s3[1] = 4096; p.out[0].s3[2] = 1;
  p.shape[0] = 1; p.shape[1] = 1024; p.shape[2] = 4096;
  std::lock_guard<std::mutex> lock(jit::g_mu);
  B200_REQUIRE(jit::load_nvrtc(), B200_ERR_UNSUPPORTED, "libnvrtc is not available");
  uint64_t total = 0;
  auto check = [&](const std::string &src, const char *name) -> int32_t {
    size_t n = 0;
    jit::compile(src, name, false, &n);
    B200_REQUIRE(n > 0, B200_ERR_CUDA, "NVRTC produced no cubin for %s", name);
    total += n;
    return B200_OK;
  };
  if ((st = check(jit::generate(ct, p, true), "b200_jit_kernel")) != B200_OK) return st;
  p.in[2].mode = kModeVec; p.in[2].s3[1] = 4096; p.in[2].s3[2] = 1;
  if ((st = check(jit::generate(ct, p, false), "b200_jit_kernel")) != B200_OK) return st;
  // the same chain without its output as a fuse-on-read tape
  ct.ops.back().dst_out = -1;
  ct.n_out = 0;
  if ((st = check(jit::gen_reduce(ct, p, B200_RED_SUM, false), "b200_jit_rows_cta")) != B200_OK) return st;
  if ((st = check(jit::gen_reduce(ct, p, B200_RED_MIN, false), "b200_jit_rows_warp")) != B200_OK) return st;
  if ((st = check(jit::gen_reduce(ct, p, B200_RED_MAX, true), "b200_jit_cols")) != B200_OK) return st;
  // every opcode of the ISA instantiated once, so an intrinsic NVRTC does not know surfaces here
  std::string all = "#include \"burn_b200.h\"\n#include \"tape_eval.cuh\"\nusing namespace b200;\n"
                    "extern \"C\" __global__ void b200_jit_allops(uint32_t *p) {\n  uint32_t a = p[0], b = p[1], c = p[2];\n";
  for (int k = 0; k < kIOpCount; ++k) all += "  a = eval_op<" + std::to_string(k) + ">(a, b, c);\n";
  all += "  p[3] = a;\n}\n";
  if ((st = check(all, "b200_jit_allops")) != B200_OK) return st;
  if (cubin_bytes_total) *cubin_bytes_total = total;
  return B200_OK;
}

extern "C" int32_t b200_jit_cache_stats(uint64_t *entries, uint64_t *evictions) {
  std::lock_guard<std::mutex> lock(b200::jit::g_mu);
  if (entries) *entries = b200::jit::g_cache.size();
  if (evictions) *evictions = b200::jit::g_evictions;
  return B200_OK;
}
