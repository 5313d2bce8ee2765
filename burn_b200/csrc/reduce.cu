// Path (b): axis / full reductions with a fuse-on-read prologue tape and a
// fuse-on-write epilogue tape.
//
// Replaces reduce_kernel_fused (crates/burn-cubecl-fusion/src/optim/reduce/
// optimization.rs:503) and the eager reduce entry points sum / reduce_dim
// (crates/burn-cubecl/src/kernel/reduce/base.rs:108-150).  Semantics follow the
// CPU oracle: keepdim outputs (crates/burn-ndarray/src/ops/macros.rs:1-60),
// mean = sum / len, argmax/argmin = first extreme with first-NaN-wins
// (crates/burn-ndarray/src/ops/base.rs:1715-1757), max/min propagate NaN
// (default max_dim = gather(argmax), crates/burn-backend/src/backend/ops/tensor.rs:1609-1614).
//
// The input is viewed as [outer, R, inner] (R = reduced axis).  Three mappings:
//   ROW_WARP  inner == 1, short rows : one warp per row, shuffle tree.
//   ROW_CTA   inner == 1, long rows  : one CTA per (row, split); splits are
//             combined by the last-arriving CTA (ticket), deterministically.
//   COL       inner  > 1             : a CTA owns TX column-vectors x TY row
//             groups; R is split across a thread-block cluster and the CTAs'
//             partials are combined through distributed shared memory.
// Roofline: HBM bandwidth; algorithmic bytes = input bytes + output bytes.
#include <cooperative_groups.h>

#include "tape_host.cuh"
#include "reduce_fast.cuh"

namespace cg = cooperative_groups;

namespace b200 {

constexpr int kRedBlock = 256;  // threads per CTA in reduce kernels

struct Acc {
  uint32_t v;  // f32 or i32 bits
  int32_t i;   // index along the reduced axis (arg kinds)
};

struct ReduceParams {
  TapeParams rd;  // fuse-on-read tape, geometry = input shape
  TapeParams wr;  // fuse-on-write tape, geometry = output shape; INPUT(0) = reduced value
  int32_t kind;
  int32_t is_int;       // lanes are i32 (else f32)
  uint32_t outer, R, inner;
  uint32_t r_vec;       // R / VEC      (ROW mappings)
  uint32_t inner_vec;   // inner / VEC  (COL mapping)
  uint32_t splits;      // ROW_CTA: CTAs per row;  COL: cluster size along R
  uint32_t rows_per_split;  // COL: rows of R per cluster rank; ROW_CTA: vectors per split
  Acc *partials;        // ROW_CTA with splits > 1
  uint32_t *tickets;
  float mean_div;       // (float)R
};

// ----------------------------------------------------------------- combine
__device__ __forceinline__ Acc acc_identity(int kind, int is_int) {
  Acc a;
  a.i = 0x7fffffff;
  switch (kind) {
    case B200_RED_SUM: case B200_RED_MEAN: case B200_RED_ANY: a.v = 0u; break;  // 0.0f == 0
    case B200_RED_PROD: a.v = is_int ? 1u : u_of(1.0f); break;
    case B200_RED_ALL: a.v = 1u; break;
    case B200_RED_MAX: case B200_RED_ARGMAX:
      a.v = is_int ? 0x80000000u : u_of(-INFINITY); break;
    case B200_RED_MIN: case B200_RED_ARGMIN:
      a.v = is_int ? 0x7fffffffu : u_of(INFINITY); break;
    case B200_RED_MAXABS: a.v = 0u; break;
    default: a.v = 0u; break;
  }
  return a;
}

// Element → accumulator domain (ANY/ALL test non-zero; MAXABS takes |x|).
__device__ __forceinline__ uint32_t acc_map(int kind, int is_int, uint32_t x) {
  switch (kind) {
    case B200_RED_ANY: case B200_RED_ALL:
      return is_int ? (x != 0u) : (f_of(x) != 0.0f);
    case B200_RED_MAXABS:
      return is_int ? (uint32_t)abs((int32_t)x) : (x & 0x7fffffffu);
    default: return x;
  }
}

// Associative, commutative combine — safe for any tree shape.
__device__ __forceinline__ Acc acc_combine(int kind, int is_int, Acc a, Acc b) {
  Acc r = a;
  switch (kind) {
    case B200_RED_SUM: case B200_RED_MEAN:
      r.v = is_int ? a.v + b.v : u_of(__fadd_rn(f_of(a.v), f_of(b.v)));
      break;
    case B200_RED_PROD:
      r.v = is_int ? a.v * b.v : u_of(__fmul_rn(f_of(a.v), f_of(b.v)));
      break;
    case B200_RED_ANY: r.v = a.v | b.v; break;
    case B200_RED_ALL: r.v = a.v & b.v; break;
    case B200_RED_MAX: case B200_RED_MAXABS: case B200_RED_MIN: {
      const bool is_min = kind == B200_RED_MIN;
      if (is_int) {
        const int32_t x = (int32_t)a.v, y = (int32_t)b.v;
        r.v = (uint32_t)(is_min ? min(x, y) : max(x, y));
      } else {
        const float x = f_of(a.v), y = f_of(b.v);
        if (x != x) r.v = a.v;
        else if (y != y) r.v = b.v;
        else r.v = u_of(is_min ? fminf(x, y) : fmaxf(x, y));
      }
      break;
    }
    case B200_RED_ARGMAX: case B200_RED_ARGMIN: {
      const bool is_min = kind == B200_RED_ARGMIN;
      bool take_b;
      if (is_int) {
        const int32_t x = (int32_t)a.v, y = (int32_t)b.v;
        take_b = is_min ? (y < x) : (y > x);
        take_b = take_b || (y == x && b.i < a.i);
      } else {
        const float x = f_of(a.v), y = f_of(b.v);
        const bool xn = x != x, yn = y != y;
        if (xn || yn) {
          take_b = yn && (!xn || b.i < a.i);
        } else {
          take_b = is_min ? (y < x) : (y > x);
          take_b = take_b || (y == x && b.i < a.i);
        }
      }
      if (take_b) r = b;
      break;
    }
    default: break;
  }
  return r;
}

__device__ __forceinline__ Acc acc_shfl_xor(Acc a, int m) {
  Acc r;
  r.v = __shfl_xor_sync(0xffffffffu, a.v, m);
  r.i = __shfl_xor_sync(0xffffffffu, a.i, m);
  return r;
}

__device__ __forceinline__ Acc warp_reduce(int kind, int is_int, Acc a) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) a = acc_combine(kind, is_int, a, acc_shfl_xor(a, m));
  return a;
}

// ----------------------------------------------------------------- tile IO
template <int VEC, int U, int RM>
__device__ __forceinline__ void load_inputs(const TapeParams &p, const SlotFile<VEC, U, kRedBlock> &slots,
                                            const uint32_t (&vidx)[U], const bool (&ok)[U]) {
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(slots.smem);
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const Coord3 c = coords3<VEC, RM>(p, ok[u] ? vidx[u] : 0u);
    for (int k = 0; k < p.n_in; ++k) {
      const OperandDesc &d = p.in[k];
      if (VEC == 4 && d.mode == kModeVec && (d.dtype == B200_F32 || d.dtype == B200_I32)) {
        if (!ok[u]) continue;
        const int64_t off = operand_offset<VEC, RM>(p, d, vidx[u], c);
        cp_async_16(smem_base + slots.private_word(k, u) * 16u, reinterpret_cast<const uint32_t *>(d.ptr) + off);
      } else {
        uint32_t r[VEC];
        if (ok[u]) {
          load_operand<VEC, RM>(p, d, vidx[u], c, r);
        } else {
#pragma unroll
          for (int j = 0; j < VEC; ++j) r[j] = 0;
        }
        slots.put(k, u, r);
      }
    }
  }
  if constexpr (VEC == 4) cp_async_wait_all();
}

// Evaluates the read tape for U vectors and leaves the values in `val`.
template <int VEC, int U, int RM>
__device__ __forceinline__ void eval_read(const ReduceParams &P, const SlotFile<VEC, U, kRedBlock> &slots,
                                          const uint32_t (&vidx)[U], const bool (&ok)[U],
                                          uint32_t (&val)[U][VEC]) {
  load_inputs<VEC, U, RM>(P.rd, slots, vidx, ok);
#pragma unroll
  for (int u = 0; u < U; ++u)
#pragma unroll
    for (int j = 0; j < VEC; ++j) val[u][j] = 0;
  run_tape<VEC, U, kRedBlock>(
      P.rd, slots, val,
      [&](int o, const uint32_t(&x)[U][VEC]) {
#pragma unroll
        for (int u = 0; u < U; ++u)
          if (ok[u]) store_operand<VEC, RM>(P.rd, P.rd.out[o], vidx[u], coords3<VEC, RM>(P.rd, vidx[u]), x[u]);
      },
      0, P.rd.n_in);
}

// Applies mean division, runs the write tape and stores VECW consecutive outputs
// starting at output vector index `ovec`.
template <int VECW, int WRM>
__device__ __forceinline__ void finalize(const ReduceParams &P, uint32_t *wr_smem, int tid, uint32_t ovec,
                                         const Acc (&a)[VECW]) {
  SlotFile<VECW, 1, kRedBlock> slots;
  slots.smem = wr_smem;
  slots.tid = tid;
  uint32_t red[VECW];
  const bool is_arg = P.kind == B200_RED_ARGMAX || P.kind == B200_RED_ARGMIN;
#pragma unroll
  for (int j = 0; j < VECW; ++j) {
    uint32_t v = is_arg ? (uint32_t)a[j].i : a[j].v;
    if (P.kind == B200_RED_MEAN)
      v = P.is_int ? (uint32_t)((int32_t)v / (int32_t)P.R) : u_of(__fdiv_rn(f_of(v), P.mean_div));
    red[j] = v;
  }
  const Coord3 c = coords3<VECW, WRM>(P.wr, ovec);
  slots.put(0, 0, red);
  for (int k = 1; k < P.wr.n_in; ++k) {
    uint32_t r[VECW];
    load_operand<VECW, WRM>(P.wr, P.wr.in[k], ovec, c, r);
    slots.put(k, 0, r);
  }
  uint32_t acc[1][VECW];
#pragma unroll
  for (int j = 0; j < VECW; ++j) acc[0][j] = red[j];
  run_tape<VECW, 1, kRedBlock>(
      P.wr, slots, acc,
      [&](int o, const uint32_t(&x)[1][VECW]) { store_operand<VECW, WRM>(P.wr, P.wr.out[o], ovec, c, x[0]); }, 0,
      P.wr.n_in);
}

// Writes the shared scalar words of both tapes; all threads must call it.
template <int VEC, int U, int VECW>
__device__ __forceinline__ void init_both_scalars(const ReduceParams &P, uint32_t *smem, uint32_t rd_words, int tid) {
  SlotFile<VEC, U, kRedBlock> rs;
  rs.smem = smem;
  rs.tid = tid;
  init_scalars<VEC, U, kRedBlock>(P.rd, rs, P.rd.n_in + P.rd.n_tmp);
  SlotFile<VECW, 1, kRedBlock> ws;
  ws.smem = smem + rd_words;
  ws.tid = tid;
  init_scalars<VECW, 1, kRedBlock>(P.wr, ws, P.wr.n_in + P.wr.n_tmp);
  __syncthreads();
}

// Folds the VEC lanes of U vectors of a ROW mapping into one Acc.
template <int VEC, int U>
__device__ __forceinline__ Acc fold_row(const ReduceParams &P, Acc a, const uint32_t (&val)[U][VEC],
                                        const uint32_t (&ridx)[U], const bool (&ok)[U]) {
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (!ok[u]) continue;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      Acc e;
      e.v = acc_map(P.kind, P.is_int, val[u][j]);
      e.i = (int32_t)(ridx[u] * VEC + j);
      a = acc_combine(P.kind, P.is_int, a, e);
    }
  }
  return a;
}

constexpr int kWarps = kRedBlock / 32;

// ----------------------------------------------------------------- ROW_WARP
template <int VEC, int U, int RM>
__global__ void __launch_bounds__(kRedBlock)
reduce_row_warp_kernel(const __grid_constant__ ReduceParams P, uint32_t rd_words) {
  extern __shared__ __align__(16) uint32_t smem[];
  SlotFile<VEC, U, kRedBlock> slots;
  slots.smem = smem;
  slots.tid = threadIdx.x;
  init_both_scalars<VEC, U, 1>(P, smem, rd_words, threadIdx.x);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t n_rows = P.outer;
  for (uint32_t row = blockIdx.x * kWarps + warp; row < n_rows; row += gridDim.x * kWarps) {
    Acc a = acc_identity(P.kind, P.is_int);
    for (uint32_t base = 0; base < P.r_vec; base += 32 * U) {
      uint32_t vidx[U], ridx[U];
      bool ok[U];
      uint32_t val[U][VEC];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        ridx[u] = base + u * 32 + lane;
        ok[u] = ridx[u] < P.r_vec;
        vidx[u] = row * P.r_vec + ridx[u];
      }
      eval_read<VEC, U, RM>(P, slots, vidx, ok, val);
      a = fold_row<VEC, U>(P, a, val, ridx, ok);
    }
    a = warp_reduce(P.kind, P.is_int, a);
    if (lane == 0) {
      const Acc one[1] = {a};
      finalize<1, kRankGeneric>(P, smem + rd_words, threadIdx.x, row, one);
    }
    __syncwarp();
  }
}

// ----------------------------------------------------------------- ROW_CTA
__device__ __forceinline__ Acc block_reduce(int kind, int is_int, Acc a, Acc *scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  a = warp_reduce(kind, is_int, a);
  __syncthreads();  // scratch reuse across calls
  if (lane == 0) scratch[warp] = a;
  __syncthreads();
  Acc r = (lane < kWarps) ? scratch[lane] : acc_identity(kind, is_int);
  r = warp_reduce(kind, is_int, r);
  return r;  // valid in every thread of warp 0 (and all warps, same data)
}

template <int VEC, int U, int RM>
__global__ void __launch_bounds__(kRedBlock)
reduce_row_cta_kernel(const __grid_constant__ ReduceParams P, uint32_t rd_words, uint32_t slot_words) {
  extern __shared__ __align__(16) uint32_t smem[];
  SlotFile<VEC, U, kRedBlock> slots;
  slots.smem = smem;
  slots.tid = threadIdx.x;
  init_both_scalars<VEC, U, 1>(P, smem, rd_words, threadIdx.x);
  Acc *scratch = reinterpret_cast<Acc *>(smem + slot_words);
  __shared__ uint32_t s_last;
  const uint32_t n_work = P.outer * P.splits;
  for (uint32_t w = blockIdx.x; w < n_work; w += gridDim.x) {
    const uint32_t row = w / P.splits, split = w - row * P.splits;
    const uint32_t begin = split * P.rows_per_split;
    const uint32_t end = min(P.r_vec, begin + P.rows_per_split);
    Acc a = acc_identity(P.kind, P.is_int);
    for (uint32_t base = begin; base < end; base += kRedBlock * U) {
      uint32_t vidx[U], ridx[U];
      bool ok[U];
      uint32_t val[U][VEC];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        ridx[u] = base + u * kRedBlock + threadIdx.x;
        ok[u] = ridx[u] < end;
        vidx[u] = row * P.r_vec + ridx[u];
      }
      eval_read<VEC, U, RM>(P, slots, vidx, ok, val);
      a = fold_row<VEC, U>(P, a, val, ridx, ok);
    }
    a = block_reduce(P.kind, P.is_int, a, scratch);
    if (P.splits == 1) {
      if (threadIdx.x == 0) {
        const Acc one[1] = {a};
        finalize<1, kRankGeneric>(P, smem + rd_words, threadIdx.x, row, one);
      }
    } else {
      if (threadIdx.x == 0) {
        P.partials[(size_t)row * P.splits + split] = a;
        __threadfence();
        const uint32_t t = atomicAdd(&P.tickets[row], 1u);
        s_last = (t == P.splits - 1) ? 1u : 0u;
      }
      __syncthreads();
      if (s_last) {
        __threadfence();
        Acc b = acc_identity(P.kind, P.is_int);
        for (uint32_t s = threadIdx.x; s < P.splits; s += kRedBlock) {
          const uint2 raw = __ldcg(reinterpret_cast<const uint2 *>(&P.partials[(size_t)row * P.splits + s]));
          Acc e;
          e.v = raw.x;
          e.i = (int32_t)raw.y;
          b = acc_combine(P.kind, P.is_int, b, e);
        }
        b = block_reduce(P.kind, P.is_int, b, scratch);
        if (threadIdx.x == 0) {
          const Acc one[1] = {b};
          finalize<1, kRankGeneric>(P, smem + rd_words, threadIdx.x, row, one);
          P.tickets[row] = 0u;
        }
      }
    }
    __syncthreads();
  }
}

// ----------------------------------------------------------------- COL
// blockDim = (TX, TY): TX column-vectors x TY row groups.  gridDim = (col tiles,
// splits) with cluster dims (1, splits, 1).
template <int VEC, int U, int RM>
__global__ void __launch_bounds__(kRedBlock)
reduce_col_kernel(const __grid_constant__ ReduceParams P, uint32_t rd_words, uint32_t slot_words) {
  extern __shared__ __align__(16) uint32_t smem[];
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  // slot file indexed by flat thread id
  SlotFile<VEC, U, kRedBlock> slots;
  slots.smem = smem;
  slots.tid = tid;
  init_both_scalars<VEC, U, VEC>(P, smem, rd_words, tid);
  Acc *scratch = reinterpret_cast<Acc *>(smem + slot_words);  // [TY][TX][VEC], then cta result [TX][VEC]

  const uint32_t TX = blockDim.x, TY = blockDim.y;
  const uint32_t tiles_per_outer = (P.inner_vec + TX - 1) / TX;
  const uint32_t o = blockIdx.x / tiles_per_outer;
  const uint32_t cvec = (blockIdx.x - o * tiles_per_outer) * TX + threadIdx.x;
  const bool col_ok = cvec < P.inner_vec;
  const uint32_t split = blockIdx.y;
  const uint32_t r_begin = split * P.rows_per_split;
  const uint32_t r_end = min(P.R, r_begin + P.rows_per_split);

  Acc a[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) a[j] = acc_identity(P.kind, P.is_int);

  for (uint32_t base = r_begin; base < r_end; base += TY * U) {
    uint32_t vidx[U], r[U];
    bool ok[U];
    uint32_t val[U][VEC];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      r[u] = base + u * TY + threadIdx.y;
      ok[u] = col_ok && r[u] < r_end;
      vidx[u] = (o * P.R + r[u]) * P.inner_vec + cvec;
    }
    eval_read<VEC, U, RM>(P, slots, vidx, ok, val);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!ok[u]) continue;
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        Acc e;
        e.v = acc_map(P.kind, P.is_int, val[u][j]);
        e.i = (int32_t)r[u];
        a[j] = acc_combine(P.kind, P.is_int, a[j], e);
      }
    }
  }

  // combine the TY row groups through shared memory (fixed order)
  __syncthreads();
#pragma unroll
  for (int j = 0; j < VEC; ++j) scratch[(threadIdx.y * TX + threadIdx.x) * VEC + j] = a[j];
  __syncthreads();
  Acc *cta_result = scratch + TY * TX * VEC;
  if (threadIdx.y == 0) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      Acc b = scratch[threadIdx.x * VEC + j];
      for (uint32_t y = 1; y < TY; ++y)
        b = acc_combine(P.kind, P.is_int, b, scratch[(y * TX + threadIdx.x) * VEC + j]);
      a[j] = b;
      cta_result[threadIdx.x * VEC + j] = b;
    }
  }

  if (P.splits > 1) {
    // cross-CTA combine over distributed shared memory
    cg::cluster_group cluster = cg::this_cluster();
    cluster.sync();
    if (cluster.block_rank() == 0 && threadIdx.y == 0) {
      for (uint32_t rk = 1; rk < P.splits; ++rk) {
        const Acc *remote = cluster.map_shared_rank(cta_result, rk);
#pragma unroll
        for (int j = 0; j < VEC; ++j)
          a[j] = acc_combine(P.kind, P.is_int, a[j], remote[threadIdx.x * VEC + j]);
      }
    }
    cluster.sync();  // keep remote smem alive until rank 0 has read it
    if (cluster.block_rank() != 0) return;
  }
  if (threadIdx.y == 0 && col_ok) finalize<VEC, kRankGeneric>(P, smem + rd_words, tid, o * P.inner_vec + cvec, a);
}

// ----------------------------------------------------------------- host planning
struct ReducePlan {
  ReduceParams P;
  int vec;
  bool col;
};

static int32_t plan_side(const b200_tape *tape, const b200_tensor *ins, int n_ins,
                         const b200_tensor *outs, int n_outs, int rank, const int64_t *shape,
                         bool first_input_virtual, CompiledTape &ct,
                         std::vector<PlannedOperand> &planned, CollapsedLayout &L, int &rm) {
  int32_t st = compile_tape(tape, n_ins + (first_input_virtual ? 1 : 0), n_outs, ct);
  if (st != B200_OK) return st;
  planned.assign(n_ins + n_outs, PlannedOperand{});
  std::vector<PlannedOperand *> all;
  for (int i = 0; i < n_ins; ++i) {
    st = broadcast_operand(ins[i], rank, shape, "reduce input", i, planned[i]);
    if (st != B200_OK) return st;
    all.push_back(&planned[i]);
  }
  for (int i = 0; i < n_outs; ++i) {
    st = broadcast_operand(outs[i], rank, shape, "reduce output", i, planned[n_ins + i]);
    if (st != B200_OK) return st;
    for (int d = 0; d < rank; ++d)
      B200_REQUIRE(outs[i].shape[d] == shape[d], B200_ERR_SHAPE,
                   "reduce output %d: dim %d is %lld, expected %lld", i, d,
                   (long long)outs[i].shape[d], (long long)shape[d]);
    all.push_back(&planned[n_ins + i]);
  }
  L = collapse_dims(rank, shape, all);
  rm = rank_mode(L, all);
  return B200_OK;
}

static void fill_side(TapeParams &tp, const std::vector<PlannedOperand> &planned, int n_ins,
                      int n_outs, const CollapsedLayout &L, int vec, bool first_input_virtual, int rm) {
  fill_geometry(tp, L, vec, rm);
  const int shift = first_input_virtual ? 1 : 0;
  if (first_input_virtual) {
    memset(&tp.in[0], 0, sizeof(OperandDesc));
    tp.in[0].mode = kModeGather;
  }
  for (int i = 0; i < n_ins; ++i) fill_desc(tp.in[i + shift], planned[i], L.rank, vec);
  for (int i = 0; i < n_outs; ++i) {
    fill_desc(tp.out[i], planned[n_ins + i], L.rank, vec);
    if (tp.out[i].mode == kModeBcast) tp.out[i].mode = kModeGather;
  }
}

static bool is_int_dtype(int32_t dt) {
  return dt == B200_I32 || dt == B200_I64 || dt == B200_BOOL || dt == B200_U8;
}

// ----------------------------------------------------------------- fast-path dispatch
template <int K>
static int32_t launch_fast(bool col, const fast::RowParams &rp, const fast::ColParams &cp, bool warp_rows,
                           unsigned grid, cudaStream_t stream) {
  if (!col) {
    if (warp_rows) fast::reduce_row_warp_fast_kernel<K><<<grid, fast::kBlock, 0, stream>>>(rp);
    else fast::reduce_row_fast_kernel<K><<<grid, fast::kBlock, 0, stream>>>(rp);
    B200_LAUNCH_CHECK();
    return B200_OK;
  }
  if (cp.tx == 0) {   // short reduced axis, many columns: thread per column-vector
    fast::reduce_col_short_kernel<K><<<grid, fast::kBlock, 0, stream>>>(cp);
    B200_LAUNCH_CHECK();
    return B200_OK;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid, cp.splits, 1);
  cfg.blockDim = dim3(cp.tx, fast::kBlock / cp.tx, 1);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = cp.partials ? 1 : cp.splits;      // partial-row mode: plain grid, no cluster combine
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cp.tx == 32) B200_CUDA(cudaLaunchKernelEx(&cfg, fast::reduce_col_fast_kernel<K, 32>, cp));
  else if (cp.tx == 16) B200_CUDA(cudaLaunchKernelEx(&cfg, fast::reduce_col_fast_kernel<K, 16>, cp));
  else B200_CUDA(cudaLaunchKernelEx(&cfg, fast::reduce_col_fast_kernel<K, 8>, cp));
  B200_LAUNCH_CHECK();
  if (cp.partials) {
    // finish: [outer, splits <= 64, inner] -> [outer, inner] with the same tiled kernel (8 column-vectors x 32 row groups
    // per CTA, no split); the mean divides by the full R.  Fixed order: replicas stay bit-identical.
    fast::ColParams fp = cp;
    fp.x = cp.partials;
    fp.R = cp.splits;
    fp.partials = nullptr;
    fp.tx = 8;
    fp.splits = 1;
    fp.rows_per_split = fp.R;
    cfg.gridDim = dim3(cp.outer * ((cp.inner4 + 7) / 8), 1, 1);
    cfg.blockDim = dim3(8, fast::kBlock / 8, 1);
    attr[0].val.clusterDim.y = 1;
    B200_CUDA(cudaLaunchKernelEx(&cfg, fast::reduce_col_fast_kernel<K, 8>, fp));
    B200_LAUNCH_CHECK();
    count_launch(1);
  }
  return B200_OK;
}

// Returns 1 when the reduction was launched on a fast path, 0 when the caller must
// use the general tape path, < 0 on error.
static int32_t try_fast_reduce(int32_t kind, int64_t outer, int64_t R, int64_t inner, const b200_tensor &in,
                               const b200_tensor &out, cudaStream_t stream, float mean_div = 0.0f) {
  int K;
  switch (kind) {
    case B200_RED_SUM: case B200_RED_MEAN: K = fast::kSum; break;
    case B200_RED_MAX: K = fast::kMax; break;
    case B200_RED_MIN: K = fast::kMin; break;
    case B200_RED_ARGMAX: K = fast::kArgMax; break;
    case B200_RED_ARGMIN: K = fast::kArgMin; break;
    default: return 0;
  }
  if (in.dtype != B200_F32 || !is_contiguous(in) || !is_contiguous(out)) return 0;
  if (((uintptr_t)in.ptr) % 16 != 0 || R == 0) return 0;
  const bool is_arg = K >= fast::kArgMax;
  if (is_arg ? !(out.dtype == B200_I32 || out.dtype == B200_I64)
             : !(out.dtype == B200_F32 || out.dtype == B200_F16 || out.dtype == B200_BF16))
    return 0;
  const bool col = inner > 1;
  if (col ? (inner % 4 != 0) : (R % 4 != 0)) return 0;
  const int sms = sm_count();
  fast::RowParams rp = {};
  fast::ColParams cp = {};
  void *ws = nullptr;
  unsigned grid = 1;
  bool warp_rows = false;
  if (!col) {
    rp.x = reinterpret_cast<const float *>(in.ptr);
    rp.out = out.ptr;
    rp.out_dtype = out.dtype;
    rp.n_rows = (uint32_t)outer;
    rp.r4 = (uint32_t)(R / 4);
    rp.mean = kind == B200_RED_MEAN;
    rp.div = mean_div != 0.0f ? mean_div : (float)R;
    rp.splits = 1;
    rp.per_split = rp.r4;
    warp_rows = rp.r4 <= 1024 && rp.n_rows >= (uint32_t)(sms * fast::kWarpsPerBlock);
    if (warp_rows) {
      grid = std::min<uint32_t>((rp.n_rows + fast::kWarpsPerBlock - 1) / fast::kWarpsPerBlock, (uint32_t)sms * 8u);
    } else {
      const uint32_t tile = fast::kBlock * 4;
      const uint32_t target = (uint32_t)sms * 8u;
      uint32_t splits = 1;
      if (rp.n_rows < target && rp.r4 > tile * 2u)
        splits = std::min<uint32_t>((target + rp.n_rows - 1) / rp.n_rows, (rp.r4 + tile - 1) / tile);
      splits = std::max(1u, std::min(splits, 2048u));
      uint32_t per = (rp.r4 + splits - 1) / splits;
      per = ((per + tile - 1) / tile) * tile;
      splits = (rp.r4 + per - 1) / per;
      rp.splits = splits;
      rp.per_split = per;
      if (splits > 1) {
        const size_t pbytes = sizeof(fast::VI) * (size_t)rp.n_rows * splits;
        const size_t tbytes = sizeof(uint32_t) * (size_t)rp.n_rows;
        B200_CUDA(cudaMallocAsync(&ws, pbytes + tbytes, stream));
        rp.partials = reinterpret_cast<fast::VI *>(ws);
        rp.tickets = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(ws) + pbytes);
        B200_CUDA(cudaMemsetAsync(rp.tickets, 0, tbytes, stream));
      }
      grid = (unsigned)std::min<uint64_t>((uint64_t)rp.n_rows * splits, (uint64_t)sms * 8u);
    }
  } else {
    cp.x = reinterpret_cast<const float *>(in.ptr);
    cp.out = out.ptr;
    cp.out_dtype = out.dtype;
    cp.outer = (uint32_t)outer;
    cp.R = (uint32_t)R;
    cp.inner4 = (uint32_t)(inner / 4);
    cp.mean = kind == B200_RED_MEAN;
    cp.div = mean_div != 0.0f ? mean_div : (float)R;
    uint32_t tx = 32;
    while (tx > 8 && cp.outer * ((cp.inner4 + tx - 1) / tx) * 8u < (uint32_t)sms * 2u && cp.R >= 1024u) tx /= 2;
    cp.tx = tx;
    const uint32_t tiles = cp.outer * ((cp.inner4 + tx - 1) / tx);
    uint32_t splits = 1;
    while (splits < 8 && tiles * splits < (uint32_t)sms * 6u && cp.R / (splits * 2) >= 64u) splits *= 2;
    // (measured, ncu launch list: [16384, 512] 16.6 us as a 128-CTA cluster launch, 8.5 + 4.1 us this way; at 256 cluster
    //  CTAs — [8192, 1024] — the two-kernel form no longer wins, hence the < sms bound)
    if (K == fast::kSum && tiles * 8u < (uint32_t)sms && cp.R >= 4096u) {
      // tall and narrow: more row splits than a cluster holds -> partial rows + the short-axis finisher
      uint32_t tx2 = 32;
      while (tx2 > 8 && cp.outer * ((cp.inner4 + tx2 - 1) / tx2) * 64u < (uint32_t)sms * 2u) tx2 /= 2;
      const uint32_t tiles2 = cp.outer * ((cp.inner4 + tx2 - 1) / tx2);
      uint32_t s2 = std::min<uint32_t>(64u, ((uint32_t)sms * 4u + tiles2 - 1) / tiles2);
      s2 = std::max(1u, std::min(s2, cp.R / 64u));
      if (s2 > 8) {
        cp.tx = tx2;
        splits = s2;
        const size_t pbytes = sizeof(float) * (size_t)cp.outer * s2 * (size_t)inner;
        B200_CUDA(cudaMallocAsync(&ws, pbytes, stream));
        cp.partials = reinterpret_cast<float *>(ws);
        cp.splits = splits;
        cp.rows_per_split = (cp.R + splits - 1) / splits;
        grid = tiles2;
      }
    }
    if (!cp.partials) {
      cp.splits = splits;
      cp.rows_per_split = (cp.R + splits - 1) / splits;
      grid = tiles;
    }
    const uint64_t colvecs = (uint64_t)cp.outer * cp.inner4;
    if (cp.R <= 64u && colvecs >= (uint64_t)sms * fast::kBlock * 2u) {
      cp.tx = 0;
      grid = (unsigned)std::min<uint64_t>((colvecs + fast::kBlock - 1) / fast::kBlock, (uint64_t)sms * 8u);
    }
  }
  int32_t st;
  switch (K) {
    case fast::kSum: st = launch_fast<fast::kSum>(col, rp, cp, warp_rows, grid, stream); break;
    case fast::kMax: st = launch_fast<fast::kMax>(col, rp, cp, warp_rows, grid, stream); break;
    case fast::kMin: st = launch_fast<fast::kMin>(col, rp, cp, warp_rows, grid, stream); break;
    case fast::kArgMax: st = launch_fast<fast::kArgMax>(col, rp, cp, warp_rows, grid, stream); break;
    default: st = launch_fast<fast::kArgMin>(col, rp, cp, warp_rows, grid, stream); break;
  }
  if (ws) cudaFreeAsync(ws, stream);
  return st == B200_OK ? 1 : st;
}


// Plain contiguous f32 reduction [outer, R, inner] -> [outer, inner] on the fast kernels (used by
// jit.cu to finish split partials).  kind MEAN divides by `mean_div` instead of R when given.
int32_t fast_reduce_f32(int32_t kind, int64_t outer, int64_t R, int64_t inner, const float *in, float *out,
                        float mean_div, cudaStream_t stream) {
  b200_tensor ti, to;
  memset(&ti, 0, sizeof(ti));
  memset(&to, 0, sizeof(to));
  ti.ptr = const_cast<float *>(in);
  to.ptr = out;
  ti.dtype = to.dtype = B200_F32;
  ti.rank = to.rank = 3;
  ti.shape[0] = to.shape[0] = outer; ti.shape[1] = R; to.shape[1] = 1; ti.shape[2] = to.shape[2] = inner;
  ti.strides[2] = to.strides[2] = 1; ti.strides[1] = inner; to.strides[1] = inner;
  ti.strides[0] = R * inner; to.strides[0] = inner;
  return try_fast_reduce(kind, outer, R, inner, ti, to, stream, mean_div);
}
int32_t jit_try_reduce(const CompiledTape &ct, const TapeParams &p, int rank_mode_, int32_t kind, int64_t outer,
                       int64_t R, int64_t inner, float *out, cudaStream_t stream);

// Core entry: the input shape is `shape` (rank dims); dims [ax_begin, ax_end) are
// reduced together (ax_end - ax_begin == 1 for an axis reduce; the whole range
// for a full reduce, which requires it to be all dims).
static int32_t reduce_impl(int32_t kind, int rank, const int64_t *shape, int ax_begin, int ax_end,
                           const b200_tape *read, const b200_tensor *inputs, int n_inputs,
                           const b200_tape *write, const b200_tensor *write_inputs,
                           int n_write_inputs, const b200_tensor *outputs, int n_outputs,
                           const int64_t *out_shape, int out_rank, int32_t value_is_int,
                           b200_stream s) {
  B200_REQUIRE(kind >= B200_RED_SUM && kind <= B200_RED_ALL, B200_ERR_INVALID, "bad reduce kind %d", kind);
  int64_t outer = 1, R = 1, inner = 1;
  for (int d = 0; d < ax_begin; ++d) outer *= shape[d];
  for (int d = ax_begin; d < ax_end; ++d) R *= shape[d];
  for (int d = ax_end; d < rank; ++d) inner *= shape[d];
  const int64_t numel = outer * R * inner;
  B200_REQUIRE(numel < (1ll << 31), B200_ERR_UNSUPPORTED, "reduce over %lld elements exceeds 2^31", (long long)numel);
  const bool is_arg = kind == B200_RED_ARGMAX || kind == B200_RED_ARGMIN;
  if (is_arg) B200_REQUIRE(R > 0, B200_ERR_SHAPE, "Cannot compute arg over an empty axis");
  if (outer * inner == 0) return B200_OK;

  if (!read && !write && n_inputs == 1 && n_outputs == 1) {
    b200_tensor in_flat = inputs[0];
    bool shapes_match = in_flat.rank == rank;
    for (int d = 0; shapes_match && d < rank; ++d) shapes_match = in_flat.shape[d] == shape[d];
    if (shapes_match) {
      const int32_t fs = try_fast_reduce(kind, outer, R, inner, in_flat, outputs[0], resolve_stream(s));
      if (fs != 0) return fs < 0 ? fs : B200_OK;
    }
  }

  ReducePlan plan;
  ReduceParams &P = plan.P;
  memset(&P, 0, sizeof(P));

  // ---- read side
  b200_tape_op mov_in = {B200_OP_MOV, (uint8_t)B200_ARG_INPUT(0), 0, 0, B200_DST_NONE, B200_DST_NONE, {0, 0}};
  b200_tape default_read = {&mov_in, 1, nullptr, 0};
  if (!read) {
    B200_REQUIRE(n_inputs == 1, B200_ERR_INVALID, "a reduce without a read tape takes exactly one input");
    read = &default_read;
  }
  // outputs of the read tape are not supported through this entry (n_out = 0)
  std::vector<PlannedOperand> rd_planned;
  CollapsedLayout rdL;
  CompiledTape rd_ct, wr_ct;
  int rm = kRankGeneric, wrm = kRankGeneric;
  int32_t st = plan_side(read, inputs, n_inputs, nullptr, 0, rank, shape, false, rd_ct, rd_planned, rdL, rm);
  if (st != B200_OK) return st;

  // ---- write side
  b200_tape_op mov_out = {B200_OP_MOV, (uint8_t)B200_ARG_INPUT(0), 0, 0, B200_DST_NONE, 0, {0, 0}};
  b200_tape default_write = {&mov_out, 1, nullptr, 0};
  if (!write) {
    B200_REQUIRE(n_outputs == 1, B200_ERR_INVALID, "a reduce without a write tape has exactly one output");
    write = &default_write;
  }
  std::vector<PlannedOperand> wr_planned;
  CollapsedLayout wrL;
  st = plan_side(write, write_inputs, n_write_inputs, outputs, n_outputs, out_rank, out_shape, true,
                 wr_ct, wr_planned, wrL, wrm);
  if (st != B200_OK) return st;

  // ---- mapping
  const bool col = inner > 1;
  int vec = 4;
  if (rdL.shape[rdL.rank - 1] % 4 != 0) vec = 1;
  if (col ? (inner % 4 != 0) : (R % 4 != 0)) vec = 1;
  int vecw = 1;
  if (col && vec == 4 && wrL.shape[wrL.rank - 1] % 4 == 0) vecw = 4;
  if (col && vec == 4 && vecw != 4) vec = 1;  // keep read/write lane counts equal in COL
  fill_side(P.rd, rd_planned, n_inputs, 0, rdL, vec, false, rm);
  fill_side(P.wr, wr_planned, n_write_inputs, n_outputs, wrL, col ? vec : 1, true, kRankGeneric);
  P.rd.n_vec = (uint32_t)(numel / vec);
  P.wr.n_vec = (uint32_t)(outer * inner / (col ? vec : 1));
  P.kind = kind;
  P.is_int = value_is_int;
  P.outer = (uint32_t)outer;
  P.R = (uint32_t)R;
  P.inner = (uint32_t)inner;
  P.r_vec = (uint32_t)(R / vec);
  P.inner_vec = (uint32_t)(inner / vec);
  P.mean_div = (float)R;
  P.splits = 1;

  cudaStream_t stream = resolve_stream(s);
  const int sms = sm_count();
  // large fuse-on-read sum/mean/max/min over linear operands into a plain f32 output: a kernel
  // specialised from the read tape (jit.cu) does tape + local reduce; everything else: interpreter
  if (write == &default_write && vec == 4 && !value_is_int && outputs[0].dtype == B200_F32 && is_contiguous(outputs[0]) &&
      (kind == B200_RED_SUM || kind == B200_RED_MEAN || kind == B200_RED_MAX || kind == B200_RED_MIN)) {
    const int32_t js = jit_try_reduce(rd_ct, P.rd, rm, kind, outer, R, inner, reinterpret_cast<float *>(outputs[0].ptr), stream);
    if (js != 0) return js < 0 ? js : B200_OK;
  }
  const int U = (vec == 4) ? 2 : 4;
  st = finalize_tape(rd_ct, U, kRedBlock, 1, P.rd);
  if (st != B200_OK) return st;
  st = finalize_tape(wr_ct, 1, kRedBlock, 1, P.wr);
  if (st != B200_OK) return st;
  auto round16 = [](size_t b) { return (b + 15) / 16 * 16; };
  const size_t rd_slots = round16(slot_file_bytes(std::max(1, P.rd.n_in + P.rd.n_tmp), P.rd.n_scalars, vec, U, kRedBlock));
  const size_t wr_slots = round16(slot_file_bytes(std::max(1, P.wr.n_in + P.wr.n_tmp), P.wr.n_scalars, col ? vec : 1, 1, kRedBlock));
  const size_t slot_bytes = rd_slots + wr_slots;
  const uint32_t rd_words = (uint32_t)(rd_slots / 4);
  const uint32_t slot_words = (uint32_t)(slot_bytes / 4);

  if (R == 0) {
    // empty axis: identity (sum → 0, prod → 1, mean → NaN like 0/0)
    // handled by running the kernels with zero iterations.
  }

  if (!col) {
    const uint32_t n_rows = P.outer;
    const bool warp_rows = (P.r_vec <= 32u * U * 4u) && n_rows >= (uint32_t)(sms * kWarps);
    if (warp_rows) {
      const size_t smem = slot_bytes;
      auto launch = [&](auto kern) -> int32_t {
        if (smem > 48 * 1024)
          B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0;
        B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kRedBlock, smem));
        per_sm = std::max(per_sm, 1);
        const uint32_t need = (n_rows + kWarps - 1) / kWarps;
        const unsigned grid = std::max(1u, std::min<uint32_t>(need, (uint32_t)(sms * per_sm)));
        kern<<<grid, kRedBlock, smem, stream>>>(P, rd_words);
        B200_LAUNCH_CHECK();
        return B200_OK;
      };
      if (vec == 4) {
        if (rm == kRankLinear) return launch(reduce_row_warp_kernel<4, 2, kRankLinear>);
        if (rm == kRank3) return launch(reduce_row_warp_kernel<4, 2, kRank3>);
        return launch(reduce_row_warp_kernel<4, 2, kRankGeneric>);
      }
      if (rm == kRankLinear) return launch(reduce_row_warp_kernel<1, 4, kRankLinear>);
      if (rm == kRank3) return launch(reduce_row_warp_kernel<1, 4, kRank3>);
      return launch(reduce_row_warp_kernel<1, 4, kRankGeneric>);
    }
    // CTA per (row, split)
    const size_t smem = slot_bytes + sizeof(Acc) * kWarps;
    const uint32_t tile = kRedBlock * U;
    uint32_t splits = 1;
    const uint32_t target = (uint32_t)sms * 4u;
    if (n_rows < target && P.r_vec > tile * 8u) {
      splits = std::min<uint32_t>((target + n_rows - 1) / n_rows, (P.r_vec + tile * 4u - 1) / (tile * 4u));
      splits = std::max(1u, std::min(splits, 1024u));
    }
    uint32_t per_split = (P.r_vec + splits - 1) / splits;
    per_split = ((per_split + tile - 1) / tile) * tile;
    if (per_split == 0) per_split = tile;
    splits = std::max(1u, (P.r_vec + per_split - 1) / per_split);
    P.splits = splits;
    P.rows_per_split = per_split;
    void *ws = nullptr;
    if (splits > 1) {
      const size_t pbytes = sizeof(Acc) * (size_t)n_rows * splits;
      const size_t tbytes = sizeof(uint32_t) * (size_t)n_rows;
      B200_CUDA(cudaMallocAsync(&ws, pbytes + tbytes, stream));
      P.partials = reinterpret_cast<Acc *>(ws);
      P.tickets = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(ws) + pbytes);
      B200_CUDA(cudaMemsetAsync(P.tickets, 0, tbytes, stream));
    }
    auto launch = [&](auto kern) -> int32_t {
      if (smem > 48 * 1024)
        B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int per_sm = 0;
      B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kRedBlock, smem));
      per_sm = std::max(per_sm, 1);
      const uint64_t work = (uint64_t)n_rows * splits;
      const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(work, (uint64_t)sms * per_sm));
      kern<<<grid, kRedBlock, smem, stream>>>(P, rd_words, slot_words);
      B200_LAUNCH_CHECK();
      return B200_OK;
    };
    if (vec == 4) {
      st = rm == kRankLinear ? launch(reduce_row_cta_kernel<4, 2, kRankLinear>)
           : rm == kRank3    ? launch(reduce_row_cta_kernel<4, 2, kRank3>)
                             : launch(reduce_row_cta_kernel<4, 2, kRankGeneric>);
    } else {
      st = rm == kRankLinear ? launch(reduce_row_cta_kernel<1, 4, kRankLinear>)
           : rm == kRank3    ? launch(reduce_row_cta_kernel<1, 4, kRank3>)
                             : launch(reduce_row_cta_kernel<1, 4, kRankGeneric>);
    }
    if (ws) cudaFreeAsync(ws, stream);
    return st;
  }

  // ---- COL mapping
  uint32_t TX = 32;
  uint32_t tiles = P.outer * ((P.inner_vec + TX - 1) / TX);
  if (tiles * 8u < (uint32_t)sms && P.inner_vec > 8) {
    TX = 8;
    tiles = P.outer * ((P.inner_vec + TX - 1) / TX);
  }
  const uint32_t TY = kRedBlock / TX;
  uint32_t splits = 1;
  const uint32_t target = (uint32_t)sms * 4u;
  while (splits < 8 && tiles * splits < target && P.R / (splits * 2) >= TY * (uint32_t)U * 2u) splits *= 2;
  P.splits = splits;
  P.rows_per_split = (P.R + splits - 1) / splits;
  const size_t scratch = sizeof(Acc) * (size_t)(TY * TX * vec + TX * vec);
  const size_t smem = slot_bytes + scratch;

  auto launch = [&](auto kern) -> int32_t {
    if (smem > 48 * 1024)
      B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(tiles, splits, 1);
    cfg.blockDim = dim3(TX, TY, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = splits;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    B200_CUDA(cudaLaunchKernelEx(&cfg, kern, P, rd_words, slot_words));
    B200_LAUNCH_CHECK();
    return B200_OK;
  };
  if (vec == 4) {
    if (rm == kRankLinear) return launch(reduce_col_kernel<4, 2, kRankLinear>);
    if (rm == kRank3) return launch(reduce_col_kernel<4, 2, kRank3>);
    return launch(reduce_col_kernel<4, 2, kRankGeneric>);
  }
  if (rm == kRankLinear) return launch(reduce_col_kernel<1, 4, kRankLinear>);
  if (rm == kRank3) return launch(reduce_col_kernel<1, 4, kRank3>);
  return launch(reduce_col_kernel<1, 4, kRankGeneric>);
}

}  // namespace b200

using namespace b200;

extern "C" int32_t b200_launch_reduce(int32_t kind, int32_t axis, int32_t rank,
                                      const int64_t *in_shape, const b200_tape *read,
                                      const b200_tensor *inputs, int32_t n_inputs,
                                      const b200_tape *write, const b200_tensor *write_inputs,
                                      int32_t n_write_inputs, const b200_tensor *outputs,
                                      int32_t n_outputs, b200_stream s) {
  B200_REQUIRE(rank >= 1 && rank <= B200_MAX_RANK, B200_ERR_INVALID, "rank %d out of range", rank);
  B200_REQUIRE(in_shape && inputs && outputs, B200_ERR_INVALID, "null argument");
  B200_REQUIRE(n_inputs >= 1 && n_outputs >= 1, B200_ERR_INVALID, "reduce needs >=1 input and output");
  B200_REQUIRE(axis >= 0 && axis < rank, B200_ERR_SHAPE, "reduce axis %d out of range for rank %d", axis, rank);
  int64_t out_shape[B200_MAX_RANK];
  for (int d = 0; d < rank; ++d) out_shape[d] = in_shape[d];
  out_shape[axis] = 1;
  // The reduced lanes are int when the value feeding the reduce is int: decided by
  // the dtype of input 0 when there is no read tape, else by the last read op.
  int32_t is_int = is_int_dtype(inputs[0].dtype);
  if (read && read->n_ops > 0) {
    const int op = read->ops[read->n_ops - 1].op;
    if (op == B200_OP_MOV || op == B200_OP_SELECT) {
      // keeps the class of its source; approximate with input 0
    } else {
      is_int = (op >= B200_OP_EQ_F && op <= B200_OP_ISINF_F) || (op >= B200_OP_ADD_I && op <= B200_OP_NOT_B) ||
               op == B200_OP_F2I || op == B200_OP_B2I || op == B200_OP_F2B || op == B200_OP_I2B;
    }
  }
  return reduce_impl(kind, rank, in_shape, axis, axis + 1, read, inputs, n_inputs, write,
                     write_inputs, n_write_inputs, outputs, n_outputs, out_shape, rank, is_int, s);
}

extern "C" int32_t b200_launch_reduce_full(int32_t kind, const b200_tensor *input,
                                           const b200_tensor *output, b200_stream s) {
  B200_REQUIRE(input && output, B200_ERR_INVALID, "null argument");
  B200_REQUIRE(input->rank >= 1 && input->rank <= B200_MAX_RANK, B200_ERR_INVALID, "bad rank");
  B200_REQUIRE(kind != B200_RED_ARGMAX && kind != B200_RED_ARGMIN, B200_ERR_UNSUPPORTED,
               "full arg reductions are expressed as reshape + axis reduce");
  B200_REQUIRE(numel_of(output->shape, output->rank) == 1, B200_ERR_SHAPE,
               "full reduce output must hold exactly one element");
  b200_tensor out1 = *output;
  out1.rank = 1;
  out1.shape[0] = 1;
  out1.strides[0] = 1;
  const int64_t one = 1;
  return reduce_impl(kind, input->rank, input->shape, 0, input->rank, nullptr, input, 1, nullptr,
                     nullptr, 0, &out1, 1, &one, 1, is_int_dtype(input->dtype), s);
}
