// Host-side planning for tape kernels: validates a b200_tape, broadcasts and
// collapses operand layouts against the reference shape, and picks the
// vector width and per-operand access mode.  This is the job of the reference's
// launch planners (InputPlanner / OutputPlanner / VectorizationPlanner,
// crates/burn-cubecl-fusion/src/engine/launch/*.rs) done once per launch.
#pragma once
#include <algorithm>
#include <cstring>
#include <vector>

#include "tape.cuh"

namespace b200 {

struct PlannedOperand {
  void *ptr;
  int32_t dtype;
  int64_t strides[kMaxDims];  // against the (uncollapsed) reference shape
};

// Copies and validates the tape into TapeParams (ops, scalars, counts).
static inline int32_t plan_tape(const b200_tape *tape, int n_in, int n_out, TapeParams &p) {
  B200_REQUIRE(tape && tape->ops, B200_ERR_INVALID, "tape is null");
  B200_REQUIRE(tape->n_ops > 0 && tape->n_ops <= B200_MAX_TAPE_OPS, B200_ERR_INVALID,
               "tape has %d ops (limit %d)", tape->n_ops, B200_MAX_TAPE_OPS);
  B200_REQUIRE(tape->n_scalars >= 0 && tape->n_scalars <= B200_MAX_TAPE_SCALARS, B200_ERR_INVALID,
               "tape has %d scalars (limit %d)", tape->n_scalars, B200_MAX_TAPE_SCALARS);
  B200_REQUIRE(n_in >= 0 && n_in <= B200_MAX_TAPE_INPUTS, B200_ERR_INVALID,
               "%d tape inputs (limit %d)", n_in, B200_MAX_TAPE_INPUTS);
  B200_REQUIRE(n_out >= 0 && n_out <= B200_MAX_TAPE_OUTPUTS, B200_ERR_INVALID,
               "%d tape outputs (limit %d)", n_out, B200_MAX_TAPE_OUTPUTS);
  int n_tmp = 0;
  for (int i = 0; i < tape->n_ops; ++i) {
    b200_tape_op op = tape->ops[i];
    B200_REQUIRE(op.op < B200_OP_COUNT, B200_ERR_INVALID, "tape op %d: bad opcode %d", i, op.op);
    const int ar = op_arity(op.op);
    const uint8_t args[3] = {op.a, op.b, op.c};
    for (int k = 0; k < ar; ++k) {
      const int kind = args[k] >> 6, idx = args[k] & 63;
      if (kind == 0) {
        B200_REQUIRE(idx == 0, B200_ERR_INVALID, "tape op %d: malformed ACC operand", i);
        B200_REQUIRE(i > 0, B200_ERR_INVALID, "tape op 0 reads ACC before any op wrote it");
      } else if (kind == 1) {
        B200_REQUIRE(idx < n_in, B200_ERR_INVALID, "tape op %d: input %d out of range (%d)", i, idx, n_in);
      } else if (kind == 2) {
        B200_REQUIRE(idx < n_tmp, B200_ERR_INVALID, "tape op %d: temp %d read before written", i, idx);
      } else {
        B200_REQUIRE(idx < tape->n_scalars, B200_ERR_INVALID, "tape op %d: scalar %d out of range", i, idx);
      }
    }
    if (op.dst_temp != B200_DST_NONE) {
      B200_REQUIRE(op.dst_temp < B200_MAX_TAPE_TEMPS, B200_ERR_INVALID, "tape op %d: temp slot %d too large", i, op.dst_temp);
      n_tmp = std::max(n_tmp, (int)op.dst_temp + 1);
    }
    if (op.dst_out != B200_DST_NONE)
      B200_REQUIRE(op.dst_out < n_out, B200_ERR_INVALID, "tape op %d: output %d out of range (%d)", i, op.dst_out, n_out);
    op.pad[0] = (uint8_t)ar;
    op.pad[1] = 0;
    p.ops[i] = op;
  }
  for (int i = 0; i < tape->n_scalars; ++i) p.scalars[i] = tape->scalars[i];
  p.n_ops = tape->n_ops;
  p.n_in = n_in;
  p.n_out = n_out;
  p.n_tmp = n_tmp;
  return B200_OK;
}

// Broadcast a descriptor against ref_shape: size-1 dims get stride 0, other
// dims must match.
static inline int32_t broadcast_operand(const b200_tensor &t, int rank, const int64_t *ref_shape,
                                        const char *what, int index, PlannedOperand &o) {
  B200_REQUIRE(t.ptr, B200_ERR_INVALID, "%s %d has a null pointer", what, index);
  B200_REQUIRE(t.rank == rank, B200_ERR_SHAPE, "%s %d has rank %d, expected %d", what, index, t.rank, rank);
  B200_REQUIRE(dtype_size(t.dtype) > 0, B200_ERR_INVALID, "%s %d has unsupported dtype %d", what, index, t.dtype);
  o.ptr = t.ptr;
  o.dtype = t.dtype;
  for (int d = 0; d < rank; ++d) {
    if (t.shape[d] == ref_shape[d]) {
      o.strides[d] = (t.shape[d] == 1) ? 0 : t.strides[d];
    } else {
      B200_REQUIRE(t.shape[d] == 1, B200_ERR_SHAPE,
                   "%s %d: dim %d is %lld, not broadcastable to %lld", what, index, d,
                   (long long)t.shape[d], (long long)ref_shape[d]);
      o.strides[d] = 0;
    }
  }
  return B200_OK;
}

struct CollapsedLayout {
  int rank;
  int64_t shape[kMaxDims];
};

// Collapses adjacent dims that are jointly contiguous for EVERY operand and
// drops size-1 dims.  strides of all operands are rewritten in place.
static inline CollapsedLayout collapse_dims(int rank, const int64_t *shape,
                                            std::vector<PlannedOperand *> &ops) {
  CollapsedLayout L;
  int64_t shp[kMaxDims];
  int r = 0;
  // drop size-1 dims
  for (int d = 0; d < rank; ++d) {
    if (shape[d] == 1) continue;
    shp[r] = shape[d];
    for (auto *o : ops) o->strides[r] = o->strides[d];
    ++r;
  }
  if (r == 0) {
    shp[0] = 1;
    for (auto *o : ops) o->strides[0] = 0;
    r = 1;
  }
  // merge d into d+1 when stride[d] == stride[d+1] * shape[d+1] for all operands
  int w = 0;  // write index of the current merged group (outer → inner)
  for (int d = 1; d < r; ++d) {
    bool mergeable = true;
    for (auto *o : ops)
      if (o->strides[w] != o->strides[d] * shp[d]) { mergeable = false; break; }
    if (mergeable) {
      shp[w] *= shp[d];
      for (auto *o : ops) o->strides[w] = o->strides[d];
    } else {
      ++w;
      shp[w] = shp[d];
      for (auto *o : ops) o->strides[w] = o->strides[d];
    }
  }
  L.rank = w + 1;
  for (int d = 0; d < L.rank; ++d) L.shape[d] = shp[d];
  return L;
}

static inline bool vec_ok(const PlannedOperand &o, int rank, int vec) {
  if (vec == 1) return false;
  if (o.strides[rank - 1] != 1) return false;
  const int es = dtype_size(o.dtype);
  if (((uintptr_t)o.ptr) % (size_t)(es * vec) != 0) return false;
  for (int d = 0; d < rank - 1; ++d)
    if (o.strides[d] % vec != 0) return false;
  return true;
}

static inline void fill_desc(OperandDesc &d, const PlannedOperand &o, int rank, int vec) {
  d.ptr = o.ptr;
  d.dtype = o.dtype;
  for (int k = 0; k < kMaxDims; ++k) d.strides[k] = k < rank ? o.strides[k] : 0;
  if (vec_ok(o, rank, vec)) d.mode = kModeVec;
  else if (o.strides[rank - 1] == 0) d.mode = kModeBcast;
  else d.mode = kModeGather;
}

// Finishes TapeParams geometry for a collapsed layout and vector width.
static inline void fill_geometry(TapeParams &p, const CollapsedLayout &L, int vec) {
  p.rank = L.rank;
  for (int d = 0; d < kMaxDims; ++d) {
    uint32_t s = d < L.rank ? (uint32_t)L.shape[d] : 1u;
    p.shape[d] = s;
    uint32_t dv = (d == L.rank - 1) ? s / (uint32_t)vec : s;
    p.div[d] = make_fastdiv(dv == 0 ? 1 : dv);
  }
}

}  // namespace b200
