// Path (a): fused elementwise chains — one kernel per fused block.
//
// Replaces ElemwiseOptimization::execute → elemwise_fuse
// (crates/burn-cubecl-fusion/src/optim/elemwise/optimization.rs:81-171) and, with
// single-op tapes, the eager kernels kernel_binop / unary_float / kernel_cmp /
// mask_fill / mask_where / cast_element
// (crates/burn-cubecl/src/kernel/{binary,unary_float,comparison}.rs,
//  kernel/mask/{mask_fill,mask_where}.rs, kernel/cast/base.rs).
//
// HBM layout: every operand is read exactly once and every output written once,
// as 128-bit accesses; vector-eligible operands go global→shared with
// cp.async (no register staging, all of a thread's loads in flight at once),
// the tape then runs out of the thread-private slot file.  Roofline: HBM
// bandwidth; algorithmic bytes = Σ operand element sizes per element.
#include "tape_host.cuh"

namespace b200 {

__device__ __forceinline__ void cp_async_16(uint32_t smem_addr, const void *g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_addr), "l"(g));
}
__device__ __forceinline__ void cp_async_8(uint32_t smem_addr, const void *g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_addr), "l"(g));
}
__device__ __forceinline__ void cp_async_4(uint32_t smem_addr, const void *g) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_addr), "l"(g));
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.wait_all;\n" ::: "memory");
}

// Expands a raw 16-byte slot word loaded by cp.async into 32-bit lanes.
__device__ __forceinline__ void expand_raw(int32_t dtype, uint32_t (&r)[4]) {
  if (dtype == B200_BF16) {
    const uint32_t x = r[0], y = r[1];
    r[0] = x << 16; r[1] = x & 0xFFFF0000u; r[2] = y << 16; r[3] = y & 0xFFFF0000u;
  } else if (dtype == B200_F16) {
    const uint32_t x = r[0], y = r[1];
    const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&x));
    const float2 f1 = __half22float2(*reinterpret_cast<const __half2 *>(&y));
    r[0] = u_of(f0.x); r[1] = u_of(f0.y); r[2] = u_of(f1.x); r[3] = u_of(f1.y);
  } else {  // BOOL / U8
    const uint32_t v = r[0];
    r[0] = v & 0xFFu; r[1] = (v >> 8) & 0xFFu; r[2] = (v >> 16) & 0xFFu; r[3] = v >> 24;
  }
}

__device__ __forceinline__ int64_t operand_offset(const OperandDesc &d, int rank,
                                                  const uint32_t (&coord)[kMaxDims]) {
  int64_t off = 0;
#pragma unroll
  for (int k = 0; k < kMaxDims; ++k)
    if (k < rank) off += (int64_t)coord[k] * d.strides[k];
  return off;
}

// Fills the input slots of one tile.  Shared with the reduce kernels.
template <int VEC, int U>
__device__ __forceinline__ void load_tile_inputs(const TapeParams &p, const SlotFile<VEC, U> &slots,
                                                 const uint32_t (&coord)[U][kMaxDims],
                                                 const bool (&ok)[U]) {
  bool needs_expand = false;
  for (int k = 0; k < p.n_in; ++k) {
    const OperandDesc &d = p.in[k];
    if (VEC == 4 && d.mode == kModeVec && d.dtype != B200_I64) {
      const int es = (d.dtype == B200_F32 || d.dtype == B200_I32) ? 4
                     : (d.dtype == B200_BF16 || d.dtype == B200_F16) ? 2 : 1;
      needs_expand |= (es != 4);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!ok[u]) continue;
        const int64_t off = operand_offset(d, p.rank, coord[u]);
        const char *g = reinterpret_cast<const char *>(d.ptr) + off * es;
        const uint32_t sa = (uint32_t)__cvta_generic_to_shared(
            slots.base + (size_t)(k * U + u) * kTapeBlock * 4);
        if (es == 4) cp_async_16(sa, g);
        else if (es == 2) cp_async_8(sa, g);
        else cp_async_4(sa, g);
      }
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        uint32_t r[VEC];
        if (ok[u]) {
          load_operand<VEC>(d, p.rank, coord[u], r);
        } else {
#pragma unroll
          for (int j = 0; j < VEC; ++j) r[j] = 0;
        }
        slots.put(k, u, r);
      }
    }
  }
  if constexpr (VEC == 4) {
    cp_async_wait_all();
    if (needs_expand) {
      for (int k = 0; k < p.n_in; ++k) {
        const OperandDesc &d = p.in[k];
        if (d.mode == kModeVec && d.dtype != B200_F32 && d.dtype != B200_I32 &&
            d.dtype != B200_I64) {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            uint32_t r[VEC];
            slots.get(k, u, r);
            expand_raw(d.dtype, r);
            slots.put(k, u, r);
          }
        }
      }
    }
  }
}

template <int VEC, int U>
__global__ void __launch_bounds__(kTapeBlock)
elemwise_tape_kernel(const __grid_constant__ TapeParams p) {
  extern __shared__ __align__(16) uint32_t smem[];
  const SlotFile<VEC, U> slots = make_slots<VEC, U>(smem, threadIdx.x);
  constexpr uint32_t kTileVecs = kTapeBlock * U;
  const uint32_t n_tiles = (p.n_vec + kTileVecs - 1) / kTileVecs;

  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    uint32_t coord[U][kMaxDims];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t v = tile * kTileVecs + u * kTapeBlock + threadIdx.x;
      ok[u] = v < p.n_vec;
      vec_coords<VEC>(p, ok[u] ? v : 0u, coord[u]);
    }
    load_tile_inputs<VEC, U>(p, slots, coord, ok);

    uint32_t acc[U][VEC];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < VEC; ++j) acc[u][j] = 0;

    run_tape<VEC, U>(p, slots, acc, [&](int o, const uint32_t(&val)[U][VEC]) {
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (ok[u]) store_operand<VEC>(p.out[o], p.rank, coord[u], val[u]);
    });
  }
}

template <int VEC, int U>
static int32_t launch_elemwise_impl(const TapeParams &p, cudaStream_t stream) {
  const int n_slots = p.n_in + p.n_tmp;
  const size_t smem = slot_file_bytes(n_slots > 0 ? n_slots : 1, VEC, U);
  B200_REQUIRE((int)smem <= max_smem_optin(), B200_ERR_UNSUPPORTED,
               "tape needs %zu B of slot file (limit %d)", smem, max_smem_optin());
  auto kern = elemwise_tape_kernel<VEC, U>;
  if (smem > 48 * 1024)
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTapeBlock, smem));
  if (per_sm < 1) per_sm = 1;
  const uint32_t tile_vecs = kTapeBlock * U;
  const uint64_t n_tiles = ((uint64_t)p.n_vec + tile_vecs - 1) / tile_vecs;
  const uint64_t resident = (uint64_t)sm_count() * per_sm;
  const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(n_tiles, resident));
  kern<<<grid, kTapeBlock, smem, stream>>>(p);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

}  // namespace b200

using namespace b200;

extern "C" int32_t b200_launch_elemwise(const b200_tape *tape, const b200_tensor *inputs,
                                        int32_t n_inputs, const b200_tensor *outputs,
                                        int32_t n_outputs, int32_t rank,
                                        const int64_t *ref_shape, b200_stream s) {
  B200_REQUIRE(rank >= 1 && rank <= B200_MAX_RANK, B200_ERR_INVALID, "rank %d out of range", rank);
  B200_REQUIRE(ref_shape, B200_ERR_INVALID, "ref_shape is null");
  B200_REQUIRE(n_outputs >= 1, B200_ERR_INVALID, "an elementwise launch needs at least one output");
  TapeParams p;
  memset(&p, 0, sizeof(p));
  int32_t st = plan_tape(tape, n_inputs, n_outputs, p);
  if (st != B200_OK) return st;

  const int64_t numel = numel_of(ref_shape, rank);
  for (int d = 0; d < rank; ++d)
    B200_REQUIRE(ref_shape[d] >= 0, B200_ERR_SHAPE, "negative dim %d", d);
  if (numel == 0) return B200_OK;  // empty tensors: nothing to launch
  B200_REQUIRE(numel < (1ll << 31), B200_ERR_UNSUPPORTED,
               "elementwise launch over %lld elements exceeds the 2^31 index range", (long long)numel);

  std::vector<PlannedOperand> planned(n_inputs + n_outputs);
  std::vector<PlannedOperand *> all;
  for (int i = 0; i < n_inputs; ++i) {
    st = broadcast_operand(inputs[i], rank, ref_shape, "input", i, planned[i]);
    if (st != B200_OK) return st;
    all.push_back(&planned[i]);
  }
  for (int i = 0; i < n_outputs; ++i) {
    const b200_tensor &o = outputs[i];
    B200_REQUIRE(o.rank == rank, B200_ERR_SHAPE, "output %d has rank %d, expected %d", i, o.rank, rank);
    for (int d = 0; d < rank; ++d)
      B200_REQUIRE(o.shape[d] == ref_shape[d], B200_ERR_SHAPE,
                   "output %d: dim %d is %lld, expected %lld", i, d, (long long)o.shape[d],
                   (long long)ref_shape[d]);
    st = broadcast_operand(o, rank, ref_shape, "output", i, planned[n_inputs + i]);
    if (st != B200_OK) return st;
    all.push_back(&planned[n_inputs + i]);
  }
  const CollapsedLayout L = collapse_dims(rank, ref_shape, all);
  const int vec = (L.shape[L.rank - 1] % 4 == 0) ? 4 : 1;
  fill_geometry(p, L, vec);
  for (int i = 0; i < n_inputs; ++i) fill_desc(p.in[i], planned[i], L.rank, vec);
  for (int i = 0; i < n_outputs; ++i) {
    fill_desc(p.out[i], planned[n_inputs + i], L.rank, vec);
    if (p.out[i].mode == kModeBcast) p.out[i].mode = kModeGather;
  }
  p.n_vec = (uint32_t)(numel / vec);

  cudaStream_t stream = resolve_stream(s);
  if (vec == 4) return launch_elemwise_impl<4, 2>(p, stream);
  return launch_elemwise_impl<1, 4>(p, stream);
}
