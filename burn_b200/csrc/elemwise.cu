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
// as 128-bit accesses.  Vector-eligible operands stream global→shared through a
// multi-stage cp.async ring (no register staging): while the tape runs on tile t
// out of the thread-private slot file, the loads of tiles t+1 … t+S-1 are in
// flight, so HBM latency is hidden without needing high occupancy.
// Roofline: HBM bandwidth; algorithmic bytes = Σ operand element sizes per element.
#include "tape_host.cuh"

namespace b200 {

constexpr int kEwBlock = 128;  // threads per CTA (vectorised kernel)
constexpr int kEwU = 4;        // 16-byte vectors per thread per tile → 16 elements
constexpr int kEw1Block = 256; // scalar fallback kernel
constexpr int kEw1U = 4;

// Issues the asynchronous loads of the vector-eligible inputs of one tile into
// ring stage `stage`.
template <int U, int BLOCK, int RM>
__device__ __forceinline__ void prefetch_tile(const TapeParams &p, const SlotFile<4, U, BLOCK> &slots,
                                              int stage, uint32_t tile) {
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(slots.smem);
  const uint32_t v0 = tile * (BLOCK * U) + threadIdx.x;
  Coord3 c[U];
#pragma unroll
  for (int u = 0; u < U; ++u) c[u] = coords3<4, RM>(p, v0 + u * BLOCK);
  for (int k = 0; k < p.n_in; ++k) {
    const OperandDesc &d = p.in[k];
    const int es = d.async_es;
    if (es == 0) continue;
    const char *base = reinterpret_cast<const char *>(d.ptr);
    const uint32_t sa = smem_base + slots.private_word(stage * p.n_in + k, 0) * 16u;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t v = v0 + u * BLOCK;
      if (v >= p.n_vec) continue;
      const char *g = base + operand_offset<4, RM>(p, d, v, c[u]) * es;
      if (es == 4) cp_async_16(sa + u * (BLOCK * 16u), g);
      else if (es == 2) cp_async_8(sa + u * (BLOCK * 16u), g);
      else cp_async_4(sa + u * (BLOCK * 16u), g);
    }
  }
}

// Vectorised kernel: VEC = 4, S-stage ring.
template <int U, int BLOCK, int RM>
__global__ void __launch_bounds__(BLOCK)
elemwise_tape_kernel_v4(const __grid_constant__ TapeParams p, int stages, uint32_t flags) {
  extern __shared__ __align__(16) uint32_t smem[];
  SlotFile<4, U, BLOCK> slots;
  slots.smem = smem;
  slots.tid = threadIdx.x;
  constexpr uint32_t kTileVecs = BLOCK * U;
  const uint32_t n_tiles = (p.n_vec + kTileVecs - 1) / kTileVecs;
  const int tmp_base = stages * p.n_in;
  const bool has_sync_inputs = flags & 1u;   // some input is not cp.async-eligible
  const bool needs_expand = flags & 2u;      // some cp.async input is 8/16-bit

  if (p.n_scalars > 0) {
    init_scalars<4, U, BLOCK>(p, slots, tmp_base + p.n_tmp);
    __syncthreads();
  }

  // prologue: fill S-1 stages
  uint32_t next = blockIdx.x;
  for (int s = 0; s < stages - 1; ++s) {
    if (next < n_tiles) prefetch_tile<U, BLOCK, RM>(p, slots, s, next);
    cp_async_commit();
    next += gridDim.x;
  }

  int stage = 0;
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    {  // keep the ring full: the stage consumed last iteration is free again
      int ps = stage + stages - 1;
      if (ps >= stages) ps -= stages;
      if (next < n_tiles) prefetch_tile<U, BLOCK, RM>(p, slots, ps, next);
      cp_async_commit();
      next += gridDim.x;
    }
    cp_async_wait_dyn(stages - 1);

    const int in_base = stage * p.n_in;
    const uint32_t v0 = tile * kTileVecs + threadIdx.x;
    if (has_sync_inputs || needs_expand) {
      for (int k = 0; k < p.n_in; ++k) {
        const OperandDesc &d = p.in[k];
        if (d.async_es == 4) continue;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const uint32_t v = v0 + u * BLOCK;
          uint32_t r[4];
          if (d.async_es != 0) {
            slots.get(in_base + k, u, r);
            expand_raw(d.dtype, r);
          } else if (v < p.n_vec) {
            load_operand<4, RM>(p, d, v, coords3<4, RM>(p, v), r);
          } else {
            r[0] = r[1] = r[2] = r[3] = 0;
          }
          slots.put(in_base + k, u, r);
        }
      }
    }

    uint32_t acc[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[u][j] = 0;

    run_tape<4, U, BLOCK>(
        p, slots, acc,
        [&](int o, const uint32_t(&val)[U][4]) {
          const OperandDesc &d = p.out[o];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const uint32_t v = v0 + u * BLOCK;
            if (v < p.n_vec) store_operand<4, RM>(p, d, v, coords3<4, RM>(p, v), val[u]);
          }
        },
        in_base, tmp_base);

    if (++stage == stages) stage = 0;
  }
  cp_async_wait_all();
}

// Scalar fallback (innermost collapsed dim not a multiple of 4): VEC = 1.
template <int U, int BLOCK, int RM>
__global__ void __launch_bounds__(BLOCK)
elemwise_tape_kernel_v1(const __grid_constant__ TapeParams p) {
  extern __shared__ __align__(16) uint32_t smem[];
  SlotFile<1, U, BLOCK> slots;
  slots.smem = smem;
  slots.tid = threadIdx.x;
  constexpr uint32_t kTileVecs = BLOCK * U;
  const uint32_t n_tiles = (p.n_vec + kTileVecs - 1) / kTileVecs;
  if (p.n_scalars > 0) {
    init_scalars<1, U, BLOCK>(p, slots, p.n_in + p.n_tmp);
    __syncthreads();
  }
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint32_t v0 = tile * kTileVecs + threadIdx.x;
    for (int k = 0; k < p.n_in; ++k) {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t v = v0 + u * BLOCK;
        uint32_t r[1] = {0u};
        if (v < p.n_vec) load_operand<1, RM>(p, p.in[k], v, coords3<1, RM>(p, v), r);
        slots.put(k, u, r);
      }
    }
    uint32_t acc[U][1];
#pragma unroll
    for (int u = 0; u < U; ++u) acc[u][0] = 0;
    run_tape<1, U, BLOCK>(
        p, slots, acc,
        [&](int o, const uint32_t(&val)[U][1]) {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const uint32_t v = v0 + u * BLOCK;
            if (v < p.n_vec) store_operand<1, RM>(p, p.out[o], v, coords3<1, RM>(p, v), val[u]);
          }
        },
        0, p.n_in);
  }
}


// ---------------------------------------------------------------------------------------
// TMA bulk-copy variant (linear layouts): the whole CTA tile of every input is one contiguous
// run of global memory, so ONE thread issues ONE `cp.async.bulk` per input per tile (SASS
// UBLKCP) with mbarrier completion, instead of 16 per-thread LDGSTS + address arithmetic per
// input.  The bytes land exactly in the thread-private slot layout: thread t's u-th vector sits
// at element (u*BLOCK + t)*4 of the tile.  8/16-bit inputs land packed and are expanded in
// place between two CTA barriers.
__device__ __forceinline__ uint32_t ew_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int U, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
elemwise_tape_kernel_bulk(const __grid_constant__ TapeParams p, uint32_t flags) {
  extern __shared__ __align__(128) uint32_t smem[];
  __shared__ __align__(8) uint64_t mbar;
  SlotFile<4, U, BLOCK> slots;
  slots.smem = smem;
  slots.tid = threadIdx.x;
  constexpr uint32_t kTileVecs = BLOCK * U;
  const uint32_t n_tiles = (p.n_vec + kTileVecs - 1) / kTileVecs;
  const int tmp_base = p.n_in;
  const bool needs_expand = flags & 2u;

  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ew_smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (p.n_scalars > 0) init_scalars<4, U, BLOCK>(p, slots, tmp_base + p.n_tmp);
  __syncthreads();

  uint32_t parity = 0;
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint32_t vec0 = tile * kTileVecs;
    const uint32_t tile_vecs = min(kTileVecs, p.n_vec - vec0);
    if (threadIdx.x == 0) {
      // generic-proxy accesses of the previous tile → async-proxy writes of this one
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      uint32_t total = 0;
      for (int k = 0; k < p.n_in; ++k) total += tile_vecs * 4u * (uint32_t)p.in[k].async_es;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ew_smem_u32(&mbar)), "r"(total) : "memory");
      for (int k = 0; k < p.n_in; ++k) {
        const uint32_t es = (uint32_t)p.in[k].async_es;
        const char *g = reinterpret_cast<const char *>(p.in[k].ptr) + (size_t)vec0 * 4u * es;
        const uint32_t dst = ew_smem_u32(smem) + (uint32_t)(k * U * BLOCK) * 16u;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(g), "r"(tile_vecs * 4u * es), "r"(ew_smem_u32(&mbar)) : "memory");
      }
    }
    // wait for the tile
    asm volatile(
        "{\n.reg .pred p;\nBULK_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra BULK_DONE;\nbra BULK_WAIT;\nBULK_DONE:\n}\n" ::"r"(ew_smem_u32(&mbar)), "r"(parity) : "memory");
    parity ^= 1;

    const uint32_t v0 = vec0 + threadIdx.x;
    if (needs_expand) {
      // packed 8/16-bit inputs: read own raw words, barrier, rewrite as 16-byte lane words
      for (int k = 0; k < p.n_in; ++k) {
        const OperandDesc &d = p.in[k];
        if (d.async_es == 4) continue;
        uint32_t raw[U][2];
        const uint8_t *base = reinterpret_cast<const uint8_t *>(smem) + (size_t)(k * U * BLOCK) * 16;
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const uint32_t e = (uint32_t)(u * BLOCK + threadIdx.x);
          if (d.async_es == 1) {
            raw[u][0] = *reinterpret_cast<const uint32_t *>(base + e * 4);
            raw[u][1] = 0;
          } else {
            const uint2 w = *reinterpret_cast<const uint2 *>(base + e * 8);
            raw[u][0] = w.x;
            raw[u][1] = w.y;
          }
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < U; ++u) {
          uint32_t r[4] = {raw[u][0], raw[u][1], 0, 0};
          expand_raw(d.dtype, r);
          slots.put(k, u, r);
        }
      }
    }

    uint32_t acc[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[u][j] = 0;
    run_tape<4, U, BLOCK>(
        p, slots, acc,
        [&](int o, const uint32_t(&val)[U][4]) {
          const OperandDesc &d = p.out[o];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const uint32_t v = v0 + u * BLOCK;
            if (v < p.n_vec) store_operand<4, kRankLinear>(p, d, v, coords3<4, kRankLinear>(p, v), val[u]);
          }
        },
        0, tmp_base);
    __syncthreads();  // every thread is done with the slots before the next tile overwrites them
  }
}

template <typename Kern, typename... Args>
static int32_t launch_persistent(Kern kern, int block, size_t smem, uint64_t n_tiles, cudaStream_t stream,
                                 Args... args) {
  if (smem > 48 * 1024)
    B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, block, smem));
  if (per_sm < 1) per_sm = 1;
  const uint64_t resident = (uint64_t)sm_count() * per_sm;
  const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(n_tiles, resident));
  kern<<<grid, block, smem, stream>>>(args...);
  B200_LAUNCH_CHECK();
  return B200_OK;
}

static int32_t launch_v4(const CompiledTape &ct, TapeParams &p, int rm, cudaStream_t stream) {
  constexpr int U = kEwU, BLOCK = kEwBlock;
  uint32_t flags = 0;
  int n_async = 0;
  for (int k = 0; k < ct.n_in; ++k) {
    const OperandDesc &d = p.in[k];
    if (d.async_es == 0) flags |= 1u;
    else {
      ++n_async;
      if (d.async_es != 4) flags |= 2u;
    }
  }
  // TMA bulk path: linear layout, every input streams as one contiguous run per tile
  // (with a single input the per-thread path is as fast and needs no CTA barrier)
  bool bulk = rm == kRankLinear && ct.n_in >= 2 && !(flags & 1u) && !getenv("B200_EW_NO_BULK");
  for (int k = 0; k < ct.n_in && bulk; ++k) {
    const OperandDesc &d = p.in[k];
    if (d.s3[2] != 1 || ((uintptr_t)d.ptr) % 16 != 0 || ((uint64_t)p.n_vec * 4u * (uint32_t)d.async_es) % 16 != 0)
      bulk = false;
  }
  if (bulk) {
    const size_t smem_b = std::max<size_t>(slot_file_bytes(ct.n_in + ct.n_tmp, (int)ct.scalars.size(), 4, U, BLOCK), 16);
    if ((int)smem_b <= max_smem_optin()) {
      int32_t stb = finalize_tape(ct, U, BLOCK, 1, p);
      if (stb != B200_OK) return stb;
      const uint64_t tiles_b = ((uint64_t)p.n_vec + BLOCK * U - 1) / (BLOCK * U);
      return launch_persistent(elemwise_tape_kernel_bulk<U, BLOCK>, BLOCK, smem_b, tiles_b, stream, p, flags);
    }
  }
  // per-thread cp.async path.  Ring depth 1 measured best on B200 (occupancy beats the ring:
  // chain 293 us at 1 stage vs 365 us at 2); B200_EW_STAGES overrides for tuning.
  int stages = 1;
  if (const char *ev = getenv("B200_EW_STAGES")) stages = std::max(1, std::min(4, atoi(ev)));
  const size_t smem_cap = 100 * 1024;
  auto bytes_for = [&](int s) { return slot_file_bytes(s * ct.n_in + ct.n_tmp, (int)ct.scalars.size(), 4, U, BLOCK); };
  while (stages > 1 && bytes_for(stages) > smem_cap) --stages;
  const size_t smem = std::max<size_t>(bytes_for(stages), 16);
  B200_REQUIRE((int)smem <= max_smem_optin(), B200_ERR_UNSUPPORTED,
               "tape needs %zu B of slot file (limit %d)", smem, max_smem_optin());
  int32_t st = finalize_tape(ct, U, BLOCK, stages, p);
  if (st != B200_OK) return st;
  const uint64_t n_tiles = ((uint64_t)p.n_vec + BLOCK * U - 1) / (BLOCK * U);
  switch (rm) {
    case kRankLinear:
      return launch_persistent(elemwise_tape_kernel_v4<U, BLOCK, kRankLinear>, BLOCK, smem, n_tiles, stream, p, stages, flags);
    case kRank3:
      return launch_persistent(elemwise_tape_kernel_v4<U, BLOCK, kRank3>, BLOCK, smem, n_tiles, stream, p, stages, flags);
    default:
      return launch_persistent(elemwise_tape_kernel_v4<U, BLOCK, kRankGeneric>, BLOCK, smem, n_tiles, stream, p, stages, flags);
  }
}

static int32_t launch_v1(const CompiledTape &ct, TapeParams &p, int rm, cudaStream_t stream) {
  constexpr int U = kEw1U, BLOCK = kEw1Block;
  const size_t smem = std::max<size_t>(
      slot_file_bytes(ct.n_in + ct.n_tmp, (int)ct.scalars.size(), 1, U, BLOCK), 16);
  int32_t st = finalize_tape(ct, U, BLOCK, 1, p);
  if (st != B200_OK) return st;
  const uint64_t n_tiles = ((uint64_t)p.n_vec + BLOCK * U - 1) / (BLOCK * U);
  switch (rm) {
    case kRankLinear:
      return launch_persistent(elemwise_tape_kernel_v1<U, BLOCK, kRankLinear>, BLOCK, smem, n_tiles, stream, p);
    case kRank3:
      return launch_persistent(elemwise_tape_kernel_v1<U, BLOCK, kRank3>, BLOCK, smem, n_tiles, stream, p);
    default:
      return launch_persistent(elemwise_tape_kernel_v1<U, BLOCK, kRankGeneric>, BLOCK, smem, n_tiles, stream, p);
  }
}

}  // namespace b200

using namespace b200;

namespace b200 {
int32_t jit_try_elemwise(const CompiledTape &ct, const TapeParams &p, int vec, int rank_mode_, cudaStream_t stream);
}

extern "C" int32_t b200_launch_elemwise(const b200_tape *tape, const b200_tensor *inputs,
                                        int32_t n_inputs, const b200_tensor *outputs,
                                        int32_t n_outputs, int32_t rank,
                                        const int64_t *ref_shape, b200_stream s) {
  B200_REQUIRE(rank >= 1 && rank <= B200_MAX_RANK, B200_ERR_INVALID, "rank %d out of range", rank);
  B200_REQUIRE(ref_shape, B200_ERR_INVALID, "ref_shape is null");
  B200_REQUIRE(n_outputs >= 1, B200_ERR_INVALID, "an elementwise launch needs at least one output");
  CompiledTape ct;
  int32_t st = compile_tape(tape, n_inputs, n_outputs, ct);
  if (st != B200_OK) return st;

  for (int d = 0; d < rank; ++d)
    B200_REQUIRE(ref_shape[d] >= 0, B200_ERR_SHAPE, "negative dim %d", d);
  const int64_t numel = numel_of(ref_shape, rank);
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
  const int rm = rank_mode(L, all);
  TapeParams p;
  memset(&p, 0, sizeof(p));
  fill_geometry(p, L, vec, rm);
  for (int i = 0; i < n_inputs; ++i) fill_desc(p.in[i], planned[i], L.rank, vec);
  for (int i = 0; i < n_outputs; ++i) {
    fill_desc(p.out[i], planned[n_inputs + i], L.rank, vec);
    if (p.out[i].mode == kModeBcast) p.out[i].mode = kModeGather;
  }
  p.n_vec = (uint32_t)(numel / vec);

  cudaStream_t stream = resolve_stream(s);
  // large linear launches: NVRTC-specialised kernel (jit.cu); everything else: the interpreter
  const int32_t jst = jit_try_elemwise(ct, p, vec, rm, stream);
  if (jst != 0) return jst < 0 ? jst : B200_OK;
  return vec == 4 ? launch_v4(ct, p, rm, stream) : launch_v1(ct, p, rm, stream);
}
