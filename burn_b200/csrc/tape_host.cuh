// Host-side planning for tape kernels: validates a b200_tape and compiles it to
// the accumulator ISA of tape.cuh, broadcasts and collapses operand layouts
// against the reference shape, and picks the vector width and per-operand access
// mode.  This is the job of the reference's launch planners (InputPlanner /
// OutputPlanner / VectorizationPlanner, crates/burn-cubecl-fusion/src/engine/launch/*.rs)
// done once per launch.
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

// Symbolic operand of a compiled op.
struct SymArg {
  int kind;  // 0 none, 1 input, 2 temp, 3 scalar
  int idx;
};

struct SymOp {
  int op;
  SymArg b, c;
  int dst_tmp, dst_out;
};

struct CompiledTape {
  std::vector<SymOp> ops;
  std::vector<uint32_t> scalars;
  int n_in = 0, n_out = 0, n_tmp = 0;
};

static inline bool is_commutative(int op) {
  switch (op) {
    case B200_OP_ADD_F: case B200_OP_MUL_F: case B200_OP_MIN_F: case B200_OP_MAX_F:
    case B200_OP_EQ_F: case B200_OP_NE_F: case B200_OP_ADD_I: case B200_OP_MUL_I:
    case B200_OP_MIN_I: case B200_OP_MAX_I: case B200_OP_AND_I: case B200_OP_OR_I:
    case B200_OP_XOR_I: case B200_OP_EQ_I: case B200_OP_NE_I: case B200_OP_AND_B:
    case B200_OP_OR_B: case B200_OP_XOR_B:
      return true;
    default: return false;
  }
}

// Mirrored comparison: OP(x, acc) == mirror(OP)(acc, x).
static inline int mirrored_cmp(int op) {
  switch (op) {
    case B200_OP_LT_F: return B200_OP_GT_F;
    case B200_OP_LE_F: return B200_OP_GE_F;
    case B200_OP_GT_F: return B200_OP_LT_F;
    case B200_OP_GE_F: return B200_OP_LE_F;
    case B200_OP_LT_I: return B200_OP_GT_I;
    case B200_OP_LE_I: return B200_OP_GE_I;
    case B200_OP_GT_I: return B200_OP_LT_I;
    case B200_OP_GE_I: return B200_OP_LE_I;
    default: return -1;
  }
}

// Validates the public tape and lowers it to acc = OP(acc, B[, C]).
static inline int32_t compile_tape(const b200_tape *tape, int n_in, int n_out, CompiledTape &ct) {
  B200_REQUIRE(tape && tape->ops, B200_ERR_INVALID, "tape is null");
  B200_REQUIRE(tape->n_ops > 0 && tape->n_ops <= B200_MAX_TAPE_OPS, B200_ERR_INVALID,
               "tape has %d ops (limit %d)", tape->n_ops, B200_MAX_TAPE_OPS);
  B200_REQUIRE(tape->n_scalars >= 0 && tape->n_scalars <= B200_MAX_TAPE_SCALARS, B200_ERR_INVALID,
               "tape has %d scalars (limit %d)", tape->n_scalars, B200_MAX_TAPE_SCALARS);
  B200_REQUIRE(n_in >= 0 && n_in <= B200_MAX_TAPE_INPUTS, B200_ERR_INVALID,
               "%d tape inputs (limit %d)", n_in, B200_MAX_TAPE_INPUTS);
  B200_REQUIRE(n_out >= 0 && n_out <= B200_MAX_TAPE_OUTPUTS, B200_ERR_INVALID,
               "%d tape outputs (limit %d)", n_out, B200_MAX_TAPE_OUTPUTS);
  ct.n_in = n_in;
  ct.n_out = n_out;
  ct.scalars.assign(tape->scalars, tape->scalars + tape->n_scalars);
  int n_tmp = 0;
  // pass 1: validate, count user temps
  for (int i = 0; i < tape->n_ops; ++i) {
    const b200_tape_op &op = tape->ops[i];
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
  }
  const int scratch = n_tmp;  // compiler-owned temp for ACC used as a non-first operand
  bool used_scratch = false;

  auto sym = [](uint8_t a) { return SymArg{a >> 6, a & 63}; };
  auto emit = [&](int op, SymArg b, SymArg c, int dt, int dout) {
    ct.ops.push_back(SymOp{op, b, c, dt, dout});
  };
  const SymArg none{0, 0};

  auto scalar_is = [&](uint8_t arg, uint32_t bits) {
    return (arg >> 6) == 3 && ct.scalars[arg & 63] == bits;
  };
  auto no_dst = [](const b200_tape_op &o) { return o.dst_temp == B200_DST_NONE && o.dst_out == B200_DST_NONE; };

  for (int i = 0; i < tape->n_ops; ++i) {
    const b200_tape_op &pop = tape->ops[i];
    int op = pop.op;
    const int ar = op_arity(op);
    SymArg a = sym(pop.a), b = ar >= 2 ? sym(pop.b) : none, c = ar >= 3 ? sym(pop.c) : none;
    const int dt = pop.dst_temp == B200_DST_NONE ? -1 : pop.dst_temp;
    const int dout = pop.dst_out == B200_DST_NONE ? -1 : pop.dst_out;

    // ---- macro-op fusion (identical arithmetic, fewer dispatches) ----
    // gelu: DIV(X, sqrt2) ERF ADD(.,1) MUL(X,.) DIV(.,2) with X re-readable (input / temp)
    if (op == B200_OP_DIV_F && i + 4 < tape->n_ops && a.kind != 0 && a.kind != 3 &&
        scalar_is(pop.b, 0x3FB504F3u) && no_dst(pop)) {
      const b200_tape_op &o1 = tape->ops[i + 1], &o2 = tape->ops[i + 2], &o3 = tape->ops[i + 3],
                         &o4 = tape->ops[i + 4];
      const bool erf_ok = o1.op == B200_OP_ERF_F && o1.a == B200_ARG_ACC && no_dst(o1);
      const bool add_ok = o2.op == B200_OP_ADD_F && no_dst(o2) &&
                          ((o2.a == B200_ARG_ACC && scalar_is(o2.b, 0x3F800000u)) ||
                           (o2.b == B200_ARG_ACC && scalar_is(o2.a, 0x3F800000u)));
      const bool mul_ok = o3.op == B200_OP_MUL_F && no_dst(o3) &&
                          ((o3.a == pop.a && o3.b == B200_ARG_ACC) || (o3.b == pop.a && o3.a == B200_ARG_ACC));
      const bool div_ok = o4.op == B200_OP_DIV_F && o4.a == B200_ARG_ACC && scalar_is(o4.b, 0x40000000u);
      if (erf_ok && add_ok && mul_ok && div_ok) {
        // X was just produced by the previous op and is not read again: keep it in ACC
        // instead of a round trip through a temp slot.
        SymArg src = a;
        if (a.kind == 2 && !ct.ops.empty() && ct.ops.back().dst_tmp == a.idx) {
          bool read_later = false;
          for (int j = i + 5; j < tape->n_ops && !read_later; ++j) {
            const b200_tape_op &oj = tape->ops[j];
            const int arj = op_arity(oj.op);
            const uint8_t args[3] = {oj.a, oj.b, oj.c};
            for (int k = 0; k < arj; ++k)
              if ((args[k] >> 6) == 2 && (args[k] & 63) == a.idx) read_later = true;
            if (oj.dst_temp == a.idx) break;  // overwritten before any further read
          }
          if (!read_later) {
            ct.ops.back().dst_tmp = -1;
            src = none;
          }
        }
        emit(kOpGelu, src, none, o4.dst_temp == B200_DST_NONE ? -1 : o4.dst_temp,
             o4.dst_out == B200_DST_NONE ? -1 : o4.dst_out);
        i += 4;
        continue;
      }
    }
    if (ar == 1) {
      if (op == B200_OP_MOV) {
        if (a.kind == 0) emit(kOpSave, none, none, dt, dout);
        else emit(kOpLoad, a, none, dt, dout);
        continue;
      }
      if (a.kind != 0) emit(kOpLoad, a, none, -1, -1);
      emit(op, none, none, dt, dout);
      continue;
    }
    // arity 2 / 3: first operand must be ACC, the others must be in memory
    if (ar == 2 && a.kind != 0 && b.kind == 0) {
      if (is_commutative(op)) std::swap(a, b);
      else if (mirrored_cmp(op) >= 0) { op = mirrored_cmp(op); std::swap(a, b); }
    }
    const bool acc_later = (ar >= 2 && b.kind == 0) || (ar >= 3 && c.kind == 0);
    if (acc_later) {
      emit(kOpSave, none, none, scratch, -1);
      used_scratch = true;
      if (b.kind == 0 && ar >= 2) b = SymArg{2, scratch};
      if (c.kind == 0 && ar >= 3) c = SymArg{2, scratch};
    }
    if (a.kind != 0) emit(kOpLoad, a, none, -1, -1);
    if (op == B200_OP_DIV_F && b.kind == 3) {
      // division by a launch constant: power of two → exact multiply; otherwise the
      // exact Markstein sequence when the divisor is safely normal
      const uint32_t bits = ct.scalars[b.idx];
      const uint32_t expo = (bits >> 23) & 0xFF, mant = bits & 0x7FFFFFu;
      if (mant == 0 && expo >= 2 && expo <= 252) {
        const uint32_t rbits = (bits & 0x80000000u) | ((254u - expo) << 23);
        int idx = -1;
        for (size_t k = 0; k < ct.scalars.size(); ++k) if (ct.scalars[k] == rbits) idx = (int)k;
        if (idx < 0 && ct.scalars.size() < B200_MAX_TAPE_SCALARS) {
          ct.scalars.push_back(rbits);
          idx = (int)ct.scalars.size() - 1;
        }
        if (idx >= 0) { op = B200_OP_MUL_F; b = SymArg{3, idx}; }
      } else if (expo >= 40 && expo <= 214 && mant != 0x7FFFFFu) {
        op = kOpDivScalar;
      }
    }
    // powf_scalar with exponent 2 or 3 (powi_scalar lowers to it: the reference gelu_backward's x^3): multiplications
    // instead of the ~40-instruction powf
    if (op == B200_OP_POW_F && b.kind == 3) {
      const uint32_t bits = ct.scalars[b.idx];
      if (bits == 0x40000000u) { op = kOpSquare; b = none; }
      else if (bits == 0x40400000u) { op = kOpCube; b = none; }
    }
    // mul-then-add with both extra operands in memory: acc = acc*B + C (two roundings)
    if (op == B200_OP_MUL_F && dt < 0 && dout < 0 && b.kind != 0 && i + 1 < tape->n_ops) {
      const b200_tape_op &o1 = tape->ops[i + 1];
      if (o1.op == B200_OP_ADD_F && ((o1.a == B200_ARG_ACC) != (o1.b == B200_ARG_ACC))) {
        const SymArg ad = sym(o1.a == B200_ARG_ACC ? o1.b : o1.a);
        emit(kOpMulAdd, b, ad, o1.dst_temp == B200_DST_NONE ? -1 : o1.dst_temp,
             o1.dst_out == B200_DST_NONE ? -1 : o1.dst_out);
        i += 1;
        continue;
      }
    }
    emit(op, b, c, dt, dout);
  }
  // temps actually referenced after lowering (peepholes may have removed some)
  int used_tmp = 0;
  for (const SymOp &o : ct.ops) {
    if (o.dst_tmp >= 0) used_tmp = std::max(used_tmp, o.dst_tmp + 1);
    if (o.b.kind == 2) used_tmp = std::max(used_tmp, o.b.idx + 1);
    if (o.c.kind == 2) used_tmp = std::max(used_tmp, o.c.idx + 1);
  }
  (void)n_tmp;
  (void)used_scratch;
  ct.n_tmp = used_tmp;
  B200_REQUIRE((int)ct.ops.size() <= kMaxIOps, B200_ERR_UNSUPPORTED,
               "tape compiles to %zu internal ops (limit %d)", ct.ops.size(), kMaxIOps);
  return B200_OK;
}

// Resolves symbolic operands to slot-file word addresses for a kernel geometry
// (U vectors per thread, BLOCK threads, `stages` input ring stages) and writes
// the program into TapeParams.
static inline int32_t finalize_tape(const CompiledTape &ct, int U, int BLOCK, int stages, TapeParams &p) {
  const int n_private = stages * ct.n_in + ct.n_tmp;
  const int64_t max_word = (int64_t)n_private * U * BLOCK + (int64_t)ct.scalars.size();
  B200_REQUIRE(max_word < 65536, B200_ERR_UNSUPPORTED, "slot file too large for 16-bit operand addresses");
  const int tmp_base = stages * ct.n_in;
  auto addr = [&](const SymArg &a, uint32_t &flags, uint32_t in_flag, uint32_t sh_flag) -> uint32_t {
    switch (a.kind) {
      case 1: flags |= in_flag; return (uint32_t)(a.idx * U * BLOCK);
      case 2: return (uint32_t)((tmp_base + a.idx) * U * BLOCK);
      case 3: flags |= sh_flag; return (uint32_t)(n_private * U * BLOCK + a.idx);
      default: return 0;
    }
  };
  for (size_t i = 0; i < ct.ops.size(); ++i) {
    const SymOp &o = ct.ops[i];
    uint32_t flags = 0;
    const uint32_t ba = addr(o.b, flags, kFlagBInput, kFlagBShared);
    const uint32_t ca = addr(o.c, flags, kFlagCInput, kFlagCShared);
    if (o.b.kind != 0) flags |= kFlagHasB;
    const uint32_t lo = (uint32_t)o.op | ((uint32_t)(o.dst_tmp < 0 ? 0xFF : o.dst_tmp) << 8) |
                        ((uint32_t)(o.dst_out < 0 ? 0xFF : o.dst_out) << 16) | (flags << 24);
    const uint32_t hi = ba | (ca << 16);
    p.iops[i] = (uint64_t)lo | ((uint64_t)hi << 32);
  }
  p.n_iops = (int32_t)ct.ops.size();
  p.n_in = ct.n_in;
  p.n_out = ct.n_out;
  p.n_tmp = ct.n_tmp;
  p.n_scalars = (int32_t)ct.scalars.size();
  for (size_t i = 0; i < ct.scalars.size(); ++i) p.scalars[i] = ct.scalars[i];
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
// `barrier_before` (optional, -1 = none) names a dim that must not be merged
// with the dim before it, and `barrier_after` one that must not be merged with
// the dim after it (reductions keep the reduced axis separate).
static inline CollapsedLayout collapse_dims(int rank, const int64_t *shape,
                                            std::vector<PlannedOperand *> &ops) {
  CollapsedLayout L;
  int64_t shp[kMaxDims];
  int r = 0;
  for (int d = 0; d < rank; ++d) {  // drop size-1 dims
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
  int w = 0;  // merge d into the current group when stride[w] == stride[d] * shape[d] for all
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

// Largest |element offset| the operand can be addressed at.
static inline int64_t max_offset(const PlannedOperand &o, int rank, const int64_t *shape) {
  int64_t m = 0;
  for (int d = 0; d < rank; ++d) m += (shape[d] - 1) * (o.strides[d] < 0 ? -o.strides[d] : o.strides[d]);
  return m;
}

// Fills the device descriptor.  For rank <= 3 strides are right-aligned into
// s3[0..2] to match TapeParams geometry; the generic path keeps all strides.
static inline void fill_desc(OperandDesc &d, const PlannedOperand &o, int rank, int vec) {
  d.ptr = o.ptr;
  d.dtype = o.dtype;
  for (int k = 0; k < kMaxDims; ++k) d.strides[k] = k < rank ? o.strides[k] : 0;
  d.s3[0] = d.s3[1] = d.s3[2] = 0;
  if (rank <= 3)
    for (int k = 0; k < rank; ++k) d.s3[3 - rank + k] = (int32_t)o.strides[k];
  if (vec_ok(o, rank, vec)) d.mode = kModeVec;
  else if (o.strides[rank - 1] == 0) d.mode = kModeBcast;
  else d.mode = kModeGather;
  d.async_es = 0;
  if (d.mode == kModeVec && o.dtype != B200_I64) d.async_es = dtype_size(o.dtype);
}

// Picks the offset arithmetic a kernel instance must use.
static inline int rank_mode(const CollapsedLayout &L, const std::vector<PlannedOperand *> &ops) {
  if (L.rank > 3) return kRankGeneric;
  for (auto *o : ops)
    if (max_offset(*o, L.rank, L.shape) >= (1ll << 31)) return kRankGeneric;
  return L.rank == 1 ? kRankLinear : kRank3;
}

// Finishes TapeParams geometry for a collapsed layout and vector width.
static inline void fill_geometry(TapeParams &p, const CollapsedLayout &L, int vec, int rm) {
  p.rank = L.rank;
  for (int d = 0; d < kMaxDims; ++d) {
    p.shape[d] = 1;
    p.div[d] = make_fastdiv(1);
  }
  const int shift = (rm != kRankGeneric && L.rank <= 3) ? 3 - L.rank : 0;
  for (int d = 0; d < L.rank; ++d) {
    const uint32_t s = (uint32_t)L.shape[d];
    p.shape[d + shift] = s;
    uint32_t dv = (d == L.rank - 1) ? s / (uint32_t)vec : s;
    p.div[d + shift] = make_fastdiv(dv == 0 ? 1 : dv);
  }
}

}  // namespace b200
