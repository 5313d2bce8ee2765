// Host-side fusion layer: a lazy operation stream, its relative form + Context, the OperationFuser state
// machines (ElementWise, Reduce, Matmul, ReduceBroadcasted), relative Optimizations with execute / to_state /
// from_state, and a plan cache — on top of the kernel ABI.
//
// Mirrors, for the hot path only, what burn-fusion + burn-cubecl-fusion do in the reference:
//   OperationIr stream     crates/burn-ir/src/operation.rs:113-142
//   relative form/Context  crates/burn-fusion/src/stream/context.rs:11-26 (Context), :56-90 (OperationConverter:
//                          tensor ids by first appearance, relative shape ids with 1 ↦ 0, scalars and slice ranges
//                          lifted out so the same graph at other sizes / values maps to the same relative trace)
//   OperationFuser         crates/burn-fusion/src/backend.rs:187-206  (fuse / finish / reset / status / properties /
//                          len / clone_dyn) — struct Fuser below, one subclass per registered fuser
//   acceptance rules       crates/burn-cubecl-fusion/src/engine/fuser.rs:76-190,292-710 (same output shape, <= 64
//                          ops, bounded bindings, Drop absorbed so intermediates stay in registers),
//                          optim/reduce/fuser.rs:103-160,221-300 (read block → one *Dim reduce → write block),
//                          optim/matmul/fuser.rs:69-151 (matmul + epilogue)
//   score                  crates/burn-cubecl-fusion/src/engine/scoring.rs:56-76: saved global reads+writes × 100 +
//                          saved launches × 10; the ready fuser with the highest score wins
//                          (crates/burn-fusion/src/search/block.rs:430-450)
//   Optimization           crates/burn-fusion/src/backend.rs:226-234: execute(context) resolves relative ids to
//                          handles, allocates (or aliases, engine/launch/output.rs:47-55) outputs, ONE launch,
//                          registers outputs, frees consumed ReadWrite handles; to_state/from_state round-trip
//   plan store             crates/burn-fusion/src/stream/store/base.rs: relative trace → list of optimizations;
//                          a hit re-executes them with a new Context and runs no fuser
#include <algorithm>
#include <atomic>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../csrc/common.cuh"
#include "burn_b200_host.h"

namespace b200h {

using b200::fail;

static std::atomic<uint64_t> g_inplace_aliases{0};

struct Buffer {
  void *ptr = nullptr;
  bool fake = false;   // plan-only streams: no memory behind it
  bool owned = true;   // false: wraps caller-owned device memory (b200h_from_device)
  ~Buffer() {
    if (ptr && !fake && owned) b200_free(ptr, nullptr);
  }
};

struct Tensor {
  std::vector<int64_t> shape, strides;
  int32_t dtype = B200_F32;
  std::shared_ptr<Buffer> buf;  // null while the producing op is still queued
  int64_t offset = 0;           // elements
};

enum class Kind : uint8_t {
  Binary, Scalar, Unary, MaskFill, MaskWhere, ReduceDim, Matmul, Drop,
  Reshape, Expand, Slice, SwapDims,  // BaseOperationIr views: metadata only
  Gather, Select                     // NumericOperationIr indexed reads: their own block
};

struct Op {
  Kind kind = Kind::Drop;
  int opcode = 0;  // b200_opcode / b200_reduce_kind
  int64_t in[3] = {-1, -1, -1};
  int64_t out = -1;
  double scalar = 0;
  int dim = 0, dim2 = 0;
  int precision = 0;
  std::vector<int64_t> range;  // Slice: starts then ends
};

static bool is_cmp(int op) {
  return (op >= B200_OP_EQ_F && op <= B200_OP_ISINF_F) || (op >= B200_OP_EQ_I && op <= B200_OP_GE_I) ||
         (op >= B200_OP_AND_B && op <= B200_OP_NOT_B) || op == B200_OP_F2B || op == B200_OP_I2B;
}
static bool is_int_op(int op) { return op >= B200_OP_ADD_I && op <= B200_OP_GE_I; }
static bool is_elemwise(Kind k) {
  return k == Kind::Binary || k == Kind::Scalar || k == Kind::Unary || k == Kind::MaskFill || k == Kind::MaskWhere;
}
static bool is_view(Kind k) { return k == Kind::Reshape || k == Kind::Expand || k == Kind::Slice || k == Kind::SwapDims; }

static std::vector<int64_t> contiguous(const std::vector<int64_t> &shape) {
  std::vector<int64_t> st(shape.size());
  int64_t acc = 1;
  for (int d = (int)shape.size() - 1; d >= 0; --d) {
    st[d] = acc;
    acc *= std::max<int64_t>(shape[d], 1);
  }
  return st;
}
static int64_t numel(const std::vector<int64_t> &shape) {
  int64_t n = 1;
  for (auto s : shape) n *= s;
  return n;
}
static bool is_contig(const std::vector<int64_t> &shape, const std::vector<int64_t> &strides) {
  int64_t expect = 1;
  for (int d = (int)shape.size() - 1; d >= 0; --d) {
    if (shape[d] != 1 && strides[d] != expect) return false;
    expect *= shape[d];
  }
  return true;
}

// ---------------------------------------------------------------- relative form
// What a fuser is allowed to look at.  Everything concrete (ids, sizes, scalar values, slice ranges) lives in the
// Context; the few concrete facts the acceptance rules need are folded into flag bits so that equal relative
// traces always lead to equal decisions.
enum : uint32_t {
  kTensorContig = 1u,   // materialised with dense row-major strides, or still pending (will be allocated dense)
  kRedLastAxis = 1u,    // ReduceDim: the reduced axis is the last one
  kRedRowOnChip = 2u,   // ReduceDim: one row fits the row-resident kernels' shared memory
  kMatmulN4 = 1u        // Matmul: output columns % 4 == 0 (fused epilogue vector width)
};
struct RelTensor {
  int32_t dtype = 0;
  uint32_t flags = 0;
  std::vector<int32_t> dims;  // relative shape ids (0 ↔ extent 1)
};
struct RelOp {
  uint8_t kind = 0;
  int32_t opcode = 0;
  int32_t in[3] = {-1, -1, -1};
  int32_t out = -1;
  int32_t scalar = -1;  // Context::scalars index
  int32_t range = -1;   // Context::ranges index
  int32_t dim = 0, dim2 = 0, precision = 0;
  uint32_t flags = 0;
};
struct Trace {
  std::vector<RelOp> ops;
  std::vector<RelTensor> tensors;
};
struct Context {
  std::vector<int64_t> tensor;  // relative tensor id → global id
  std::vector<int64_t> dim;     // relative shape id → extent
  std::vector<double> scalar;
  std::vector<std::vector<int64_t>> ranges;
};

struct ByteWriter {
  std::string s;
  template <class T> void put(const T &v) { s.append(reinterpret_cast<const char *>(&v), sizeof(T)); }
  template <class T> void vec(const std::vector<T> &v) {
    put<uint32_t>((uint32_t)v.size());
    if (!v.empty()) s.append(reinterpret_cast<const char *>(v.data()), v.size() * sizeof(T));
  }
};
struct ByteReader {
  const char *p, *end;
  bool ok = true;
  template <class T> T get() {
    T v{};
    if (end - p < (ptrdiff_t)sizeof(T)) { ok = false; return v; }
    memcpy(&v, p, sizeof(T));
    p += sizeof(T);
    return v;
  }
  template <class T> std::vector<T> vec() {
    const uint32_t n = get<uint32_t>();
    std::vector<T> v;
    if (!ok || (uint64_t)(end - p) < (uint64_t)n * sizeof(T)) { ok = false; return v; }
    v.resize(n);
    if (n) memcpy(v.data(), p, n * sizeof(T));
    p += n * sizeof(T);
    return v;
  }
};

static void write_trace(ByteWriter &w, const Trace &t) {
  w.put<uint32_t>((uint32_t)t.ops.size());
  for (const RelOp &o : t.ops) {
    w.put(o.kind); w.put(o.opcode);
    for (int k = 0; k < 3; ++k) w.put(o.in[k]);
    w.put(o.out); w.put(o.scalar); w.put(o.range); w.put(o.dim); w.put(o.dim2); w.put(o.precision); w.put(o.flags);
  }
  w.put<uint32_t>((uint32_t)t.tensors.size());
  for (const RelTensor &x : t.tensors) {
    w.put(x.dtype); w.put(x.flags);
    w.vec(x.dims);
  }
}
static bool read_trace(ByteReader &r, Trace &t) {
  const uint32_t n = r.get<uint32_t>();
  if (!r.ok || n > (1u << 20)) return false;
  t.ops.resize(n);
  for (RelOp &o : t.ops) {
    o.kind = r.get<uint8_t>(); o.opcode = r.get<int32_t>();
    for (int k = 0; k < 3; ++k) o.in[k] = r.get<int32_t>();
    o.out = r.get<int32_t>(); o.scalar = r.get<int32_t>(); o.range = r.get<int32_t>(); o.dim = r.get<int32_t>();
    o.dim2 = r.get<int32_t>(); o.precision = r.get<int32_t>(); o.flags = r.get<uint32_t>();
  }
  const uint32_t m = r.get<uint32_t>();
  if (!r.ok || m > (1u << 20)) return false;
  t.tensors.resize(m);
  for (RelTensor &x : t.tensors) {
    x.dtype = r.get<int32_t>(); x.flags = r.get<uint32_t>();
    x.dims = r.vec<int32_t>();
  }
  return r.ok;
}
static std::string trace_key(const Trace &t) {
  ByteWriter w;
  write_trace(w, t);
  return std::move(w.s);
}

// ---------------------------------------------------------------- tape assembly (relative ids)
struct ScalarSrc {
  int32_t ctx = 0;      // Context::scalars index
  int32_t as_int = 0;   // integer opcode: the value is an i32, not f32 bits
  bool operator==(const ScalarSrc &o) const { return ctx == o.ctx && as_int == o.as_int; }
};
struct TapeBuild {
  std::vector<b200_tape_op> ops;
  std::vector<ScalarSrc> scalars;
  std::vector<int32_t> inputs;   // relative tensor ids, in INPUT(k) order
  std::vector<int32_t> outputs;  // relative tensor ids written, in out-index order
  bool ok = false;
};

static uint32_t f32_bits(double v) {
  float f = (float)v;
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}

// Builds the public tape for a run of elementwise ops of `tr`.  `virtual_in0` (>= 0) is a tensor that must map to
// INPUT(0) without being a real global input (reduced value / accumulator).
static TapeBuild build_tape(const Trace &tr, const std::vector<int> &ops, const std::unordered_set<int32_t> &dropped,
                            int32_t virtual_in0, bool virtual_needs_output, int max_outputs) {
  TapeBuild tb;
  std::unordered_map<int32_t, int> producer;  // tensor id -> op index in block
  for (size_t i = 0; i < ops.size(); ++i) producer[tr.ops[ops[i]].out] = (int)i;
  std::vector<int> last_use(ops.size(), -1);
  std::vector<bool> far_use(ops.size(), false);
  for (size_t i = 0; i < ops.size(); ++i)
    for (int k = 0; k < 3; ++k) {
      auto it = producer.find(tr.ops[ops[i]].in[k]);
      if (tr.ops[ops[i]].in[k] >= 0 && it != producer.end() && it->second < (int)i) {
        last_use[it->second] = (int)i;
        if ((int)i > it->second + 1) far_use[it->second] = true;
      }
    }
  if (virtual_in0 >= 0) tb.inputs.push_back(virtual_in0);
  auto input_index = [&](int32_t id) -> int {
    for (size_t k = 0; k < tb.inputs.size(); ++k)
      if (tb.inputs[k] == id) return (int)k;
    tb.inputs.push_back(id);
    return (int)tb.inputs.size() - 1;
  };
  auto scalar_index = [&](int32_t ctx, bool as_int) -> int {
    const ScalarSrc s = {ctx, as_int ? 1 : 0};
    for (size_t k = 0; k < tb.scalars.size(); ++k)
      if (tb.scalars[k] == s) return (int)k;
    tb.scalars.push_back(s);
    return (int)tb.scalars.size() - 1;
  };
  if (virtual_in0 >= 0 && virtual_needs_output) {
    b200_tape_op t = {B200_OP_MOV, (uint8_t)B200_ARG_INPUT(0), 0, 0, B200_DST_NONE, (uint8_t)tb.outputs.size(), {0, 0}};
    tb.outputs.push_back(virtual_in0);
    tb.ops.push_back(t);
  }
  std::vector<int> temp_of(ops.size(), -1);
  std::vector<int> temp_free_at(B200_MAX_TAPE_TEMPS, -1);  // op index after which the slot is free
  for (size_t i = 0; i < ops.size(); ++i) {
    const RelOp &o = tr.ops[ops[i]];
    auto arg = [&](int32_t id) -> int {
      auto it = producer.find(id);
      if (it != producer.end() && it->second < (int)i) {
        const int j = it->second;
        if (j == (int)i - 1) return B200_ARG_ACC;
        return B200_ARG_TEMP(temp_of[j]);
      }
      return B200_ARG_INPUT(input_index(id));
    };
    b200_tape_op t;
    memset(&t, 0, sizeof(t));
    t.dst_temp = B200_DST_NONE;
    t.dst_out = B200_DST_NONE;
    switch ((Kind)o.kind) {
      case Kind::Binary:
        t.op = (uint8_t)o.opcode; t.a = (uint8_t)arg(o.in[0]); t.b = (uint8_t)arg(o.in[1]);
        break;
      case Kind::Scalar:
        t.op = (uint8_t)o.opcode; t.a = (uint8_t)arg(o.in[0]);
        t.b = (uint8_t)B200_ARG_SCALAR(scalar_index(o.scalar, is_int_op(o.opcode)));
        break;
      case Kind::Unary:
        t.op = (uint8_t)o.opcode; t.a = (uint8_t)arg(o.in[0]);
        break;
      case Kind::MaskFill:  // ConditionalAssign: cond ? scalar : x
        t.op = B200_OP_SELECT; t.a = (uint8_t)arg(o.in[0]);
        t.b = (uint8_t)B200_ARG_SCALAR(scalar_index(o.scalar, false)); t.c = (uint8_t)arg(o.in[1]);
        break;
      case Kind::MaskWhere:
        t.op = B200_OP_SELECT; t.a = (uint8_t)arg(o.in[0]); t.b = (uint8_t)arg(o.in[2]); t.c = (uint8_t)arg(o.in[1]);
        break;
      default: return tb;
    }
    if (far_use[i]) {  // a value used later than the next op needs a temp slot
      int slot = -1;
      for (int s = 0; s < B200_MAX_TAPE_TEMPS; ++s)
        if (temp_free_at[s] < (int)i) { slot = s; break; }
      if (slot < 0) return tb;
      temp_of[i] = slot;
      temp_free_at[slot] = last_use[i];
      t.dst_temp = (uint8_t)slot;
    }
    if (!dropped.count(o.out)) {
      if ((int)tb.outputs.size() >= max_outputs) return tb;
      t.dst_out = (uint8_t)tb.outputs.size();
      tb.outputs.push_back(o.out);
    }
    tb.ops.push_back(t);
  }
  if (tb.ops.empty() || (int)tb.ops.size() > B200_MAX_TAPE_OPS || (int)tb.inputs.size() > B200_MAX_TAPE_INPUTS ||
      (int)tb.scalars.size() > B200_MAX_TAPE_SCALARS)
    return tb;
  tb.ok = true;
  return tb;
}

static void put_tape(ByteWriter &w, const TapeBuild &t) {
  w.put<uint8_t>(t.ok ? 1 : 0);
  w.vec(t.ops); w.vec(t.scalars); w.vec(t.inputs); w.vec(t.outputs);
}
static void get_tape(ByteReader &r, TapeBuild &t) {
  t.ok = r.get<uint8_t>() != 0;
  t.ops = r.vec<b200_tape_op>(); t.scalars = r.vec<ScalarSrc>(); t.inputs = r.vec<int32_t>(); t.outputs = r.vec<int32_t>();
}

// ---------------------------------------------------------------- Optimization (relative ids only)
enum : int { kBlockDrop = -1 };  // internal: a leading OperationIr::Drop, not logged
struct Optimization {
  int32_t kind = B200H_BLOCK_EAGER;
  int32_t consumed = 0;  // queue entries (ops + drops)
  int32_t n_ops = 0;     // IR ops excluding drops
  uint64_t score = 0;
  Trace trace;           // the relative operations this optimization stands for (its identity)
  std::vector<int32_t> read_ops, write_ops;  // indices into trace.ops
  int32_t core = -1;                         // the reduce / matmul / view / gather op
  int32_t row_in = -1, row_out = -1, row_kind = 0, row_gamma = -1, row_beta = -1, row_eps = -1;
  std::vector<int32_t> dropped;              // relative tensors consumed inside the block
  TapeBuild read_tape, write_tape;

  const char *name() const {
    switch (kind) {
      case B200H_BLOCK_ELEMWISE: return "ElementWise";
      case B200H_BLOCK_REDUCE: return "Reduce";
      case B200H_BLOCK_MATMUL: return "Matmul";
      case B200H_BLOCK_ROWNORM: return "ReduceBroadcasted";
      case B200H_BLOCK_VIEW: return "View";
      case kBlockDrop: return "Drop";
      default: return "Unfused";
    }
  }
  std::string to_state() const {
    ByteWriter w;
    w.put<uint32_t>(0x504f3242u);  // "B2OP"
    w.put<uint32_t>(1);            // format version
    w.put(kind); w.put(consumed); w.put(n_ops); w.put(score);
    write_trace(w, trace);
    w.vec(read_ops); w.vec(write_ops);
    w.put(core); w.put(row_in); w.put(row_out); w.put(row_kind); w.put(row_gamma); w.put(row_beta); w.put(row_eps);
    w.vec(dropped);
    put_tape(w, read_tape);
    put_tape(w, write_tape);
    return std::move(w.s);
  }
  static std::shared_ptr<Optimization> from_state(const void *buf, uint64_t len) {
    ByteReader r{(const char *)buf, (const char *)buf + len};
    if (r.get<uint32_t>() != 0x504f3242u || r.get<uint32_t>() != 1) return nullptr;
    auto o = std::make_shared<Optimization>();
    o->kind = r.get<int32_t>(); o->consumed = r.get<int32_t>(); o->n_ops = r.get<int32_t>(); o->score = r.get<uint64_t>();
    if (!read_trace(r, o->trace)) return nullptr;
    o->read_ops = r.vec<int32_t>(); o->write_ops = r.vec<int32_t>();
    o->core = r.get<int32_t>(); o->row_in = r.get<int32_t>(); o->row_out = r.get<int32_t>(); o->row_kind = r.get<int32_t>();
    o->row_gamma = r.get<int32_t>(); o->row_beta = r.get<int32_t>(); o->row_eps = r.get<int32_t>();
    o->dropped = r.vec<int32_t>();
    get_tape(r, o->read_tape);
    get_tape(r, o->write_tape);
    if (!r.ok || r.p != r.end || o->consumed != (int32_t)o->trace.ops.size()) return nullptr;
    // every index must stay inside the trace it came with
    const int32_t nt = (int32_t)o->trace.tensors.size(), no = (int32_t)o->trace.ops.size();
    auto tin = [&](int32_t v) { return v >= -1 && v < nt; };
    auto oin = [&](int32_t v) { return v >= -1 && v < no; };
    bool ok = oin(o->core) && tin(o->row_in) && tin(o->row_out) && tin(o->row_gamma) && tin(o->row_beta);
    for (auto v : o->read_ops) ok = ok && v >= 0 && oin(v);
    for (auto v : o->write_ops) ok = ok && v >= 0 && oin(v);
    for (auto v : o->dropped) ok = ok && v >= 0 && tin(v);
    for (const TapeBuild *t : {&o->read_tape, &o->write_tape}) {
      for (auto v : t->inputs) ok = ok && v >= 0 && tin(v);
      for (auto v : t->outputs) ok = ok && v >= 0 && tin(v);
    }
    int32_t n_scalars = 0;
    for (const RelOp &op : o->trace.ops) {
      ok = ok && tin(op.out);
      for (int k = 0; k < 3; ++k) ok = ok && tin(op.in[k]);
      if (op.scalar >= 0) {
        ok = ok && op.scalar <= n_scalars;
        n_scalars = std::max(n_scalars, op.scalar + 1);
      }
    }
    for (const TapeBuild *t : {&o->read_tape, &o->write_tape})
      for (const ScalarSrc &sc : t->scalars) ok = ok && sc.ctx >= 0 && sc.ctx < n_scalars;
    switch (o->kind) {
      case B200H_BLOCK_ELEMWISE: ok = ok && !o->read_ops.empty() && o->read_tape.ok; break;
      case B200H_BLOCK_REDUCE:
        ok = ok && o->core >= 0 && (o->read_ops.empty() || o->read_tape.ok) && (o->write_ops.empty() || (o->write_tape.ok && !o->write_tape.inputs.empty()));
        break;
      case B200H_BLOCK_MATMUL:
        ok = ok && o->core >= 0 && (o->write_ops.empty() || (o->write_tape.ok && o->write_tape.outputs.size() == 1 && !o->write_tape.inputs.empty()));
        break;
      case B200H_BLOCK_ROWNORM:
        ok = ok && o->row_in >= 0 && o->row_out >= 0 && o->row_kind >= 0 && o->row_kind <= 2 &&
             (o->row_kind != 2 || (o->row_gamma >= 0 && o->row_eps >= 0 && o->row_eps < n_scalars));
        break;
      case B200H_BLOCK_VIEW: case B200H_BLOCK_EAGER: case kBlockDrop: ok = ok && o->consumed == 1; break;
      default: ok = false;
    }
    return ok ? o : nullptr;
  }
};

// scoring.rs:56-76 — I/O the unfused ops would have done minus what the fused block does, plus saved launches.
static uint64_t score_of(const Trace &tr, const std::vector<int> &ops, size_t fused_io) {
  size_t unfused = 0;
  for (int i : ops) {
    const RelOp &o = tr.ops[i];
    for (int k = 0; k < 3; ++k) unfused += o.in[k] >= 0;
    unfused += 1;
  }
  const uint64_t io = unfused > fused_io ? (uint64_t)(unfused - fused_io) * 100 : 0;
  return io + (ops.empty() ? 0 : (uint64_t)(ops.size() - 1) * 10);
}

// The prefix of `tr` an optimization consuming `consumed` entries stands for (first-appearance numbering makes the
// tensor table prefix-stable).
static Trace trace_prefix(const Trace &tr, int consumed) {
  Trace p;
  p.ops.assign(tr.ops.begin(), tr.ops.begin() + consumed);
  int32_t hi = -1;
  for (const RelOp &o : p.ops) {
    hi = std::max(hi, o.out);
    for (int k = 0; k < 3; ++k) hi = std::max(hi, o.in[k]);
  }
  p.tensors.assign(tr.tensors.begin(), tr.tensors.begin() + (hi + 1));
  return p;
}

// ---------------------------------------------------------------- fusers
struct Fuser {
  int32_t status = B200H_FUSER_OPEN;
  int fed = 0;  // entries offered so far (= index of the next one)
  Optimization best;  // last valid state; n_ops == 0 → nothing to finish
  virtual ~Fuser() {}
  virtual int block_kind() const = 0;
  virtual void fuse(const Trace &tr, int idx) = 0;
  virtual void reset() = 0;
  virtual std::unique_ptr<Fuser> clone() const = 0;
  bool ready() const { return best.n_ops > 0; }
  uint64_t score() const { return best.score; }
  int len() const { return best.n_ops; }
  void offer(const Trace &tr, int idx) {
    if (status == B200H_FUSER_OPEN) fuse(tr, idx);
    fed = idx + 1;
  }
  std::shared_ptr<Optimization> finish(const Trace &tr) {
    if (!ready()) return nullptr;
    auto o = std::make_shared<Optimization>(best);
    o->trace = trace_prefix(tr, o->consumed);
    return o;
  }
  void base_reset() { status = B200H_FUSER_OPEN; fed = 0; best = Optimization(); }
};

static std::vector<int> concat(const std::vector<int32_t> &a, int core, const std::vector<int32_t> &b) {
  std::vector<int> v(a.begin(), a.end());
  if (core >= 0) v.push_back(core);
  v.insert(v.end(), b.begin(), b.end());
  return v;
}

// ---- ElementWise: longest run of same-shape elementwise ops (drops absorbed)
struct ElemwiseFuser : Fuser {
  std::vector<int> ops;
  std::unordered_set<int32_t> dropped;
  std::vector<int32_t> shape;
  int entries = 0;
  int block_kind() const override { return B200H_BLOCK_ELEMWISE; }
  void reset() override { base_reset(); ops.clear(); dropped.clear(); shape.clear(); entries = 0; }
  std::unique_ptr<Fuser> clone() const override { return std::unique_ptr<Fuser>(new ElemwiseFuser(*this)); }
  void snapshot(const Trace &tr, const TapeBuild &tb) {
    best = Optimization();
    best.kind = B200H_BLOCK_ELEMWISE;
    best.consumed = entries;
    best.n_ops = (int)ops.size();
    best.read_ops.assign(ops.begin(), ops.end());
    best.dropped.assign(dropped.begin(), dropped.end());
    std::sort(best.dropped.begin(), best.dropped.end());
    best.read_tape = tb;
    best.score = score_of(tr, ops, tb.inputs.size() + tb.outputs.size());
  }
  void fuse(const Trace &tr, int idx) override {
    const RelOp &o = tr.ops[idx];
    if ((Kind)o.kind == Kind::Drop) {
      if (ops.empty()) { status = B200H_FUSER_CLOSED; return; }  // leading drops are applied by the stream
      dropped.insert(o.in[0]);
      ++entries;
      TapeBuild tb = build_tape(tr, ops, dropped, -1, false, B200_MAX_TAPE_OUTPUTS);
      if (tb.ok) snapshot(tr, tb);
      return;
    }
    if (!is_elemwise((Kind)o.kind) || (int)ops.size() >= B200_MAX_TAPE_OPS ||
        (!ops.empty() && tr.tensors[o.out].dims != shape)) {  // output_is_compatible: same shape only
      status = B200H_FUSER_CLOSED;
      return;
    }
    ops.push_back(idx);
    TapeBuild tb = build_tape(tr, ops, dropped, -1, false, B200_MAX_TAPE_OUTPUTS);
    if (!tb.ok) {
      ops.pop_back();
      status = B200H_FUSER_CLOSED;
      return;
    }
    if (ops.size() == 1) shape = tr.tensors[o.out].dims;
    ++entries;
    snapshot(tr, tb);
  }
};

// ---- Reduce: [read block] → ReduceDim → [write block]
struct ReduceFuser : Fuser {
  std::vector<int> read_ops, write_ops;
  std::unordered_set<int32_t> dropped;
  std::vector<int32_t> in_shape;
  int core = -1, entries = 0;
  int block_kind() const override { return B200H_BLOCK_REDUCE; }
  void reset() override { base_reset(); read_ops.clear(); write_ops.clear(); dropped.clear(); in_shape.clear(); core = -1; entries = 0; }
  std::unique_ptr<Fuser> clone() const override { return std::unique_ptr<Fuser>(new ReduceFuser(*this)); }
  void evaluate(const Trace &tr) {
    Optimization c;
    c.kind = B200H_BLOCK_REDUCE;
    c.consumed = entries;
    c.n_ops = (int)(read_ops.size() + 1 + write_ops.size());
    c.core = core;
    const RelOp &red = tr.ops[core];
    size_t io = 0;
    if (!read_ops.empty()) {
      // nothing produced by the read block may need materialising (the kernel has no read outputs)
      for (int i : read_ops)
        if (!dropped.count(tr.ops[i].out)) return;
      c.read_tape = build_tape(tr, read_ops, dropped, -1, false, 0);
      if (!c.read_tape.ok) return;
      io += c.read_tape.inputs.size();
    } else {
      io += 1;
    }
    if (!write_ops.empty()) {
      c.write_tape = build_tape(tr, write_ops, dropped, red.out, !dropped.count(red.out), B200_MAX_TAPE_OUTPUTS);
      if (!c.write_tape.ok) return;
      io += c.write_tape.inputs.size() - 1 + c.write_tape.outputs.size();
    } else {
      io += 1;
    }
    c.read_ops.assign(read_ops.begin(), read_ops.end());
    c.write_ops.assign(write_ops.begin(), write_ops.end());
    c.dropped.assign(dropped.begin(), dropped.end());
    std::sort(c.dropped.begin(), c.dropped.end());
    c.score = score_of(tr, concat(c.read_ops, core, c.write_ops), io);
    best = c;
  }
  void fuse(const Trace &tr, int idx) override {
    const RelOp &o = tr.ops[idx];
    const Kind k = (Kind)o.kind;
    if (core < 0) {  // read phase
      if (k == Kind::Drop) {
        if (read_ops.empty()) { status = B200H_FUSER_CLOSED; return; }
        dropped.insert(o.in[0]);
        ++entries;
      } else if (k == Kind::ReduceDim) {
        // the reduced value must be the last read op's result and the read shape the reduce input shape
        if (!read_ops.empty() && (tr.ops[read_ops.back()].out != o.in[0] || tr.tensors[o.in[0]].dims != in_shape)) {
          status = B200H_FUSER_CLOSED;
          return;
        }
        core = idx;
        ++entries;
        evaluate(tr);
      } else if (is_elemwise(k) && (int)read_ops.size() < B200_MAX_TAPE_OPS &&
                 (read_ops.empty() || tr.tensors[o.out].dims == in_shape)) {
        if (read_ops.empty()) in_shape = tr.tensors[o.out].dims;
        read_ops.push_back(idx);
        ++entries;
      } else {
        status = B200H_FUSER_CLOSED;
      }
      return;
    }
    if (k == Kind::Drop) {
      dropped.insert(o.in[0]);
      ++entries;
      evaluate(tr);
      return;
    }
    bool ok = is_elemwise(k) && tr.tensors[o.out].dims == tr.tensors[tr.ops[core].out].dims &&
              (int)write_ops.size() < B200_MAX_TAPE_OPS;
    // write-block inputs must be readable at the output shape: reject reads of read-block values
    for (int j = 0; ok && j < 3; ++j)
      for (int r : read_ops)
        if (o.in[j] >= 0 && o.in[j] == tr.ops[r].out) ok = false;
    if (!ok) { status = B200H_FUSER_CLOSED; return; }
    write_ops.push_back(idx);
    ++entries;
    evaluate(tr);
  }
};

// ---- Matmul: Matmul → [epilogue block on the output shape]
struct MatmulFuser : Fuser {
  std::vector<int> write_ops;
  std::unordered_set<int32_t> dropped;
  int core = -1, entries = 0;
  int block_kind() const override { return B200H_BLOCK_MATMUL; }
  void reset() override { base_reset(); write_ops.clear(); dropped.clear(); core = -1; entries = 0; }
  std::unique_ptr<Fuser> clone() const override { return std::unique_ptr<Fuser>(new MatmulFuser(*this)); }
  void fuse(const Trace &tr, int idx) override {
    const RelOp &o = tr.ops[idx];
    const Kind k = (Kind)o.kind;
    if (core < 0) {
      if (k != Kind::Matmul) { status = B200H_FUSER_CLOSED; return; }
      core = idx;
      entries = 1;
      best = Optimization();
      best.kind = B200H_BLOCK_MATMUL;
      best.consumed = 1;
      best.n_ops = 1;
      best.core = idx;
      best.score = 0;
      if (!(o.flags & kMatmulN4)) status = B200H_FUSER_CLOSED;  // the fused epilogue needs N % 4 == 0
      return;
    }
    const RelOp &mm = tr.ops[core];
    if (k == Kind::Drop) {
      dropped.insert(o.in[0]);
    } else {
      if (!is_elemwise(k) || tr.tensors[o.out].dims != tr.tensors[mm.out].dims) { status = B200H_FUSER_CLOSED; return; }
      write_ops.push_back(idx);
    }
    ++entries;
    if (write_ops.empty()) {  // drops of the operands directly after the product ride along
      best.consumed = entries;
      best.dropped.assign(dropped.begin(), dropped.end());
      std::sort(best.dropped.begin(), best.dropped.end());
      return;
    }
    if (!dropped.count(mm.out)) return;  // the raw product would have to be written too
    TapeBuild tb = build_tape(tr, write_ops, dropped, mm.out, false, 1);
    if (!tb.ok || tb.outputs.size() != 1) return;
    Optimization c;
    c.kind = B200H_BLOCK_MATMUL;
    c.consumed = entries;
    c.n_ops = 1 + (int)write_ops.size();
    c.core = core;
    c.write_ops.assign(write_ops.begin(), write_ops.end());
    c.dropped.assign(dropped.begin(), dropped.end());
    std::sort(c.dropped.begin(), c.dropped.end());
    c.write_tape = tb;
    c.score = score_of(tr, concat({}, core, c.write_ops), 2 + (tb.inputs.size() - 1) + 1);
    best = c;
  }
};

// ---- ReduceBroadcasted: the softmax / log_softmax / layer_norm chains the default ops record
//   softmax     = max_dim, sub, exp, sum_dim, div          (activation.rs:250-256)
//   log_softmax = max_dim, sub, exp, sum_dim, log, sub     (activation.rs:271-276)
//   layer_norm  = mean_dim, sub, mul, mean_dim, add_scalar, sqrt, div, mul(gamma) [, add(beta)]  (modules/base.rs:846-877)
// along the last axis, every intermediate dropped inside the window → one row-resident kernel
// (crates/burn-cubecl-fusion/src/optim/reduce_broadcasted/ fuses the same shape of work).
struct RowNormFuser : Fuser {
  std::vector<int> ops;
  std::unordered_set<int32_t> drops;
  int entries = 0;
  int block_kind() const override { return B200H_BLOCK_ROWNORM; }
  void reset() override { base_reset(); ops.clear(); drops.clear(); entries = 0; }
  std::unique_ptr<Fuser> clone() const override { return std::unique_ptr<Fuser>(new RowNormFuser(*this)); }

  // does op #i of the chain fit pattern p (0 softmax, 1 log_softmax, 2 layer_norm)?
  bool fits(const Trace &tr, int p, int i) const {
    const RelOp &o = tr.ops[ops[i]];
    const Kind k = (Kind)o.kind;
    const int32_t x = tr.ops[ops[0]].in[0];
    auto out = [&](int j) { return tr.ops[ops[j]].out; };
    auto bin = [&](int opc, int32_t a, int32_t b) { return k == Kind::Binary && o.opcode == opc && o.in[0] == a && o.in[1] == b; };
    auto un = [&](int opc, int32_t a) { return k == Kind::Unary && o.opcode == opc && o.in[0] == a; };
    auto red = [&](int kind, int32_t a) {
      return k == Kind::ReduceDim && o.opcode == kind && o.in[0] == a && (o.flags & kRedLastAxis) && (o.flags & kRedRowOnChip);
    };
    auto vec_of_row = [&](int32_t id) {  // [1, …, 1, R] dense f32
      const RelTensor &t = tr.tensors[id], &xt = tr.tensors[x];
      if (t.dtype != B200_F32 || !(t.flags & kTensorContig) || t.dims.empty() || t.dims.back() != xt.dims.back()) return false;
      for (size_t d = 0; d + 1 < t.dims.size(); ++d)
        if (t.dims[d] != 0) return false;
      return true;
    };
    if (i == 0) {
      const RelTensor &xt = tr.tensors[x];
      if (xt.dtype != B200_F32 || !(xt.flags & kTensorContig) || xt.dims.empty()) return false;
      return red(p == 2 ? B200_RED_MEAN : B200_RED_MAX, x);
    }
    if (i == 1) return bin(B200_OP_SUB_F, x, out(0));
    if (p < 2) {
      switch (i) {
        case 2: return un(B200_OP_EXP_F, out(1));
        case 3: return red(B200_RED_SUM, out(2));
        case 4: return p == 0 ? bin(B200_OP_DIV_F, out(2), out(3)) : un(B200_OP_LOG_F, out(3));
        case 5: return p == 1 && bin(B200_OP_SUB_F, out(1), out(4));
        default: return false;
      }
    }
    switch (i) {
      case 2: return bin(B200_OP_MUL_F, out(1), out(1));
      case 3: return red(B200_RED_MEAN, out(2));
      case 4: return k == Kind::Scalar && o.opcode == B200_OP_ADD_F && o.in[0] == out(3);
      case 5: return un(B200_OP_SQRT_F, out(4));
      case 6: return bin(B200_OP_DIV_F, out(1), out(5));
      case 7: return k == Kind::Binary && o.opcode == B200_OP_MUL_F && o.in[0] == out(6) && vec_of_row(o.in[1]);
      case 8: return k == Kind::Binary && o.opcode == B200_OP_ADD_F && o.in[0] == out(7) && vec_of_row(o.in[1]);
      default: return false;
    }
  }
  void try_complete(const Trace &tr, int p) {
    const int n = (int)ops.size();
    const bool complete = (p == 0 && n == 5) || (p == 1 && n == 6) || (p == 2 && (n == 8 || n == 9));
    if (!complete) return;
    const int32_t x = tr.ops[ops[0]].in[0];
    for (int i = 0; i + 1 < n; ++i)
      if (!drops.count(tr.ops[ops[i]].out)) return;  // an intermediate is still wanted: leave it to the other fusers
    Optimization c;
    c.kind = B200H_BLOCK_ROWNORM;
    c.consumed = entries;
    c.n_ops = n;
    c.row_kind = p;
    c.row_in = x;
    c.row_out = tr.ops[ops[n - 1]].out;
    if (p == 2) {
      c.row_gamma = tr.ops[ops[7]].in[1];
      c.row_eps = tr.ops[ops[4]].scalar;
      if (n == 9) c.row_beta = tr.ops[ops[8]].in[1];
    }
    c.read_ops.assign(ops.begin(), ops.end());
    c.dropped.assign(drops.begin(), drops.end());
    std::sort(c.dropped.begin(), c.dropped.end());
    c.score = score_of(tr, ops, 1 + (c.row_gamma >= 0) + (c.row_beta >= 0) + 1);
    best = c;
  }
  void fuse(const Trace &tr, int idx) override {
    const RelOp &o = tr.ops[idx];
    if ((Kind)o.kind == Kind::Drop) {
      if (ops.empty()) { status = B200H_FUSER_CLOSED; return; }
      // only the chain's own values (and its input) may be released inside the block
      bool own = o.in[0] == tr.ops[ops[0]].in[0];
      for (int i : ops) own = own || tr.ops[i].out == o.in[0];
      if (!own) { status = B200H_FUSER_CLOSED; return; }
      drops.insert(o.in[0]);
      ++entries;
      for (int p = 0; p < 3; ++p)
        if (alive(tr, p)) try_complete(tr, p);
      return;
    }
    if (ops.size() >= 9) { status = B200H_FUSER_CLOSED; return; }
    ops.push_back(idx);
    bool any = false;
    for (int p = 0; p < 3; ++p) any = any || alive(tr, p);
    if (!any) {
      ops.pop_back();
      status = B200H_FUSER_CLOSED;
      return;
    }
    ++entries;
    for (int p = 0; p < 3; ++p)
      if (alive(tr, p)) try_complete(tr, p);
  }
  bool alive(const Trace &tr, int p) const {
    const int limit = p == 0 ? 5 : p == 1 ? 6 : 9;
    if ((int)ops.size() > limit) return false;
    for (int i = 0; i < (int)ops.size(); ++i)
      if (!fits(tr, p, i)) return false;
    return true;
  }
};

static std::unique_ptr<Fuser> make_fuser(int kind) {
  switch (kind) {
    case B200H_BLOCK_ELEMWISE: return std::unique_ptr<Fuser>(new ElemwiseFuser());
    case B200H_BLOCK_REDUCE: return std::unique_ptr<Fuser>(new ReduceFuser());
    case B200H_BLOCK_MATMUL: return std::unique_ptr<Fuser>(new MatmulFuser());
    case B200H_BLOCK_ROWNORM: return std::unique_ptr<Fuser>(new RowNormFuser());
    default: return nullptr;
  }
}

// ---------------------------------------------------------------- the stream
constexpr int kMismatch = -1000;      // internal: the queue head is not what the optimization was built from
constexpr size_t kExploreWindow = 192;  // entries a fuser can be offered (64 ops + their drops, with slack)
constexpr int64_t kRowSmemLimit = 227 * 1024 - 1024;

struct Stream {
  bool plan_only = false;
  int64_t next_id = 1;
  std::unordered_map<int64_t, Tensor> tensors;
  std::deque<Op> queue;
  std::vector<b200h_block_info> log;
  std::unordered_map<std::string, std::vector<std::shared_ptr<Optimization>>> plans;
  uint64_t hits = 0, misses = 0;

  Tensor *get(int64_t id) {
    auto it = tensors.find(id);
    return it == tensors.end() ? nullptr : &it->second;
  }
  int64_t add_pending(const std::vector<int64_t> &shape, int32_t dtype) {
    Tensor t;
    t.shape = shape;
    t.strides = contiguous(shape);
    t.dtype = dtype;
    const int64_t id = next_id++;
    tensors[id] = std::move(t);
    return id;
  }

  // ---- OperationConverter: the first `count` queue entries in relative form + the Context that binds them back
  void relativise(size_t count, Trace &tr, Context &cx) {
    std::unordered_map<int64_t, int32_t> t2r;
    std::unordered_map<int64_t, int32_t> d2r;
    std::unordered_map<uint64_t, int32_t> s2r;
    d2r[1] = 0;  // global 1 is always shape id 0 (context.rs:70-72)
    cx.dim.push_back(1);
    auto rel_dim = [&](int64_t v) {
      auto it = d2r.find(v);
      if (it != d2r.end()) return it->second;
      const int32_t r = (int32_t)cx.dim.size();
      d2r[v] = r;
      cx.dim.push_back(v);
      return r;
    };
    auto rel_tensor = [&](int64_t id) -> int32_t {
      if (id < 0) return -1;
      auto it = t2r.find(id);
      if (it != t2r.end()) return it->second;
      const int32_t r = (int32_t)tr.tensors.size();
      t2r[id] = r;
      cx.tensor.push_back(id);
      RelTensor rt;
      if (const Tensor *t = get(id)) {
        rt.dtype = t->dtype;
        for (auto v : t->shape) rt.dims.push_back(rel_dim(v));
        if (!t->buf || is_contig(t->shape, t->strides)) rt.flags |= kTensorContig;
      }
      tr.tensors.push_back(std::move(rt));
      return r;
    };
    tr.ops.reserve(count);
    for (size_t q = 0; q < count; ++q) {
      const Op &o = queue[q];
      RelOp r;
      r.kind = (uint8_t)o.kind;
      r.opcode = o.opcode;
      for (int k = 0; k < 3; ++k) r.in[k] = rel_tensor(o.in[k]);
      r.out = rel_tensor(o.out);
      r.dim = o.dim;
      r.dim2 = o.dim2;
      r.precision = o.precision;
      if (o.kind == Kind::Scalar || o.kind == Kind::MaskFill) {
        // scalars are numbered by first appearance of their VALUE (as extents are): the relative form records
        // which constants coincide, not what they are, so a fused kernel binds each distinct constant once
        uint64_t bits;
        memcpy(&bits, &o.scalar, 8);
        auto it = s2r.find(bits);
        if (it == s2r.end()) {
          it = s2r.emplace(bits, (int32_t)cx.scalar.size()).first;
          cx.scalar.push_back(o.scalar);
        }
        r.scalar = it->second;
      }
      if (o.kind == Kind::Slice) {
        r.range = (int32_t)cx.ranges.size();
        cx.ranges.push_back(o.range);
      }
      if (o.kind == Kind::ReduceDim) {
        if (const Tensor *t = get(o.in[0])) {
          if (o.dim == (int)t->shape.size() - 1) r.flags |= kRedLastAxis;
          if (t->shape[o.dim] * 4 <= kRowSmemLimit) r.flags |= kRedRowOnChip;
        }
      }
      if (o.kind == Kind::Matmul) {
        if (const Tensor *t = get(o.out))
          if (t->shape.back() % 4 == 0) r.flags |= kMatmulN4;
      }
      tr.ops.push_back(r);
    }
  }

  // ---- descriptors
  int32_t desc_of(int64_t id, b200_tensor &d) {
    Tensor *t = get(id);
    B200_REQUIRE(t, B200_ERR_INVALID, "unknown tensor id %lld", (long long)id);
    B200_REQUIRE(t->buf || plan_only, B200_ERR_INVALID, "tensor %lld has no storage yet", (long long)id);
    memset(&d, 0, sizeof(d));
    d.ptr = t->buf && t->buf->ptr ? (char *)t->buf->ptr + t->offset * b200::dtype_size(t->dtype) : nullptr;
    d.dtype = t->dtype;
    d.rank = (int32_t)t->shape.size();
    for (size_t i = 0; i < t->shape.size(); ++i) {
      d.shape[i] = t->shape[i];
      d.strides[i] = t->strides[i];
    }
    return B200_OK;
  }
  int32_t materialise(int64_t id) {
    Tensor *t = get(id);
    B200_REQUIRE(t, B200_ERR_INVALID, "unknown tensor id %lld", (long long)id);
    if (t->buf) return B200_OK;
    auto b = std::make_shared<Buffer>();
    if (plan_only) {
      b->fake = true;
    } else {
      int32_t st = b200_alloc(&b->ptr, (uint64_t)std::max<int64_t>(numel(t->shape), 1) * b200::dtype_size(t->dtype), nullptr);
      if (st != B200_OK) return st;
    }
    t->buf = b;
    t->strides = contiguous(t->shape);
    t->offset = 0;
    return B200_OK;
  }
  // HandleOutput::Alias (engine/launch/output.rs:47-55): an output may be written over an input the block consumes
  // (ReadWrite: dropped inside the block, no other owner of the buffer) when both have the same dense layout.
  bool try_alias(int64_t out_id, const std::vector<int64_t> &in_ids, const std::unordered_set<int64_t> &consumed,
                 std::unordered_set<const Buffer *> &taken) {
    Tensor *o = get(out_id);
    if (!o || o->buf) return false;
    for (int64_t iid : in_ids) {
      Tensor *t = get(iid);
      if (!t || !t->buf || !consumed.count(iid) || !t->buf->owned || t->buf.use_count() != 1 || taken.count(t->buf.get())) continue;
      if (t->shape != o->shape || !is_contig(t->shape, t->strides) ||
          b200::dtype_size(t->dtype) != b200::dtype_size(o->dtype))
        continue;
      bool zero_stride = false;
      for (size_t d = 0; d < t->shape.size(); ++d) zero_stride = zero_stride || (t->shape[d] > 1 && t->strides[d] == 0);
      if (zero_stride) continue;
      taken.insert(t->buf.get());
      o->buf = t->buf;
      o->offset = t->offset;
      o->strides = contiguous(o->shape);
      return true;
    }
    return false;
  }

  struct BoundTape {
    std::vector<uint32_t> scalars;
    b200_tape tape;
  };
  static void bind_tape(const TapeBuild &tb, const Context &cx, BoundTape &bt) {
    bt.scalars.clear();
    for (const ScalarSrc &s : tb.scalars)
      bt.scalars.push_back(s.as_int ? (uint32_t)(int32_t)cx.scalar[s.ctx] : f32_bits(cx.scalar[s.ctx]));
    bt.tape.ops = tb.ops.data();
    bt.tape.n_ops = (int32_t)tb.ops.size();
    bt.tape.scalars = bt.scalars.empty() ? nullptr : bt.scalars.data();
    bt.tape.n_scalars = (int32_t)bt.scalars.size();
  }
  int32_t gather_descs(const std::vector<int32_t> &rel, size_t skip, const Context &cx, std::vector<b200_tensor> &out) {
    out.clear();
    for (size_t i = skip; i < rel.size(); ++i) {
      b200_tensor d;
      int32_t st = desc_of(cx.tensor[rel[i]], d);
      if (st != B200_OK) return st;
      out.push_back(d);
    }
    return B200_OK;
  }

  // ---- views (metadata only; a reshape of a strided tensor copies first, like into_contiguous + reshape)
  int32_t apply_view(const Op &o, int *launches) {
    Tensor *src = get(o.in[0]);
    Tensor *dst = get(o.out);
    B200_REQUIRE(src && dst && (src->buf || plan_only), B200_ERR_INVALID, "view of a tensor without storage");
    if (!src->buf) { int32_t st = materialise(o.in[0]); if (st != B200_OK) return st; src = get(o.in[0]); }
    const size_t r = src->shape.size();
    Tensor v = *src;
    switch (o.kind) {
      case Kind::SwapDims:
        std::swap(v.shape[o.dim], v.shape[o.dim2]);
        std::swap(v.strides[o.dim], v.strides[o.dim2]);
        break;
      case Kind::Expand: {
        const size_t nr = dst->shape.size();
        std::vector<int64_t> st(nr, 0);
        for (size_t d = 0; d < nr; ++d) {
          const ptrdiff_t sd = (ptrdiff_t)d - (ptrdiff_t)(nr - r);
          if (sd >= 0 && src->shape[sd] == dst->shape[d] && src->shape[sd] != 1) st[d] = src->strides[sd];
        }
        v.shape = dst->shape;
        v.strides = st;
        break;
      }
      case Kind::Slice:
        for (size_t d = 0; d < r; ++d) {
          v.offset += o.range[d] * src->strides[d];
          v.shape[d] = o.range[r + d] - o.range[d];
        }
        break;
      case Kind::Reshape:
        if (!is_contig(src->shape, src->strides)) {
          Tensor c;
          c.shape = src->shape;
          c.strides = contiguous(src->shape);
          c.dtype = src->dtype;
          c.buf = std::make_shared<Buffer>();
          if (plan_only) {
            c.buf->fake = true;
          } else {
            int32_t st = b200_alloc(&c.buf->ptr, (uint64_t)std::max<int64_t>(numel(c.shape), 1) * b200::dtype_size(c.dtype), nullptr);
            if (st != B200_OK) return st;
            b200_tensor sd, dd;
            if ((st = desc_of(o.in[0], sd)) != B200_OK) return st;
            dd = sd;
            dd.ptr = c.buf->ptr;
            for (size_t d = 0; d < r; ++d) dd.strides[d] = c.strides[d];
            if ((st = b200_launch_copy(&sd, &dd, nullptr)) != B200_OK) return st;
          }
          if (launches) ++*launches;
          v = c;
        }
        v.shape = dst->shape;
        v.strides = contiguous(dst->shape);
        break;
      default: return fail(B200_ERR_INVALID, "not a view");
    }
    *dst = std::move(v);
    return B200_OK;
  }

  // ---- Optimization::execute on the queue head with a fresh Context
  int32_t run(const Optimization &c, bool from_cache) {
    if ((size_t)c.consumed > queue.size()) return kMismatch;
    Trace tr;
    Context cx;
    relativise((size_t)c.consumed, tr, cx);
    if (trace_key(tr) != trace_key(c.trace)) return kMismatch;
    auto G = [&](int32_t rel) { return cx.tensor[rel]; };
    b200h_block_info info;
    memset(&info, 0, sizeof(info));
    info.kind = c.kind;
    info.n_ops = c.n_ops;
    info.from_cache = from_cache ? 1 : 0;
    info.score = c.score;
    int32_t st = B200_OK;
    std::vector<b200_tensor> ins, outs, wins;
    std::unordered_set<int64_t> consumed_ids;
    for (auto r : c.dropped) consumed_ids.insert(G(r));
    const uint64_t before = b200_launch_count();
    int extra_launches = 0;
    BoundTape rt, wt;
    if (c.kind == kBlockDrop) {
      tensors.erase(queue[0].in[0]);
      queue.pop_front();
      return B200_OK;
    } else if (c.kind == B200H_BLOCK_VIEW) {
      if ((st = apply_view(queue[0], &extra_launches)) != B200_OK) return st;
      info.n_inputs = 1;
      info.n_outputs = 1;
    } else if (c.kind == B200H_BLOCK_EAGER) {
      const Op &o = queue[0];
      if ((st = materialise(o.out)) != B200_OK) return st;
      b200_tensor x, idx, out;
      if ((st = desc_of(o.in[0], x)) != B200_OK || (st = desc_of(o.in[1], idx)) != B200_OK || (st = desc_of(o.out, out)) != B200_OK)
        return st;
      info.n_inputs = 2;
      info.n_outputs = 1;
      if (!plan_only)
        st = o.kind == Kind::Gather ? b200_launch_gather(o.dim, &x, &idx, &out, nullptr) : b200_launch_select(o.dim, &x, &idx, &out, nullptr);
    } else if (c.kind == B200H_BLOCK_ELEMWISE) {
      std::vector<int64_t> in_ids;
      for (auto r : c.read_tape.inputs) in_ids.push_back(G(r));
      std::unordered_set<const Buffer *> taken;
      for (auto r : c.read_tape.outputs) {
        if (try_alias(G(r), in_ids, consumed_ids, taken)) {
          ++info.aliased;
          ++g_inplace_aliases;
        } else if ((st = materialise(G(r))) != B200_OK) {
          return st;
        }
      }
      if ((st = gather_descs(c.read_tape.inputs, 0, cx, ins)) != B200_OK) return st;
      if ((st = gather_descs(c.read_tape.outputs, 0, cx, outs)) != B200_OK) return st;
      info.n_inputs = (int)ins.size();
      info.n_outputs = (int)outs.size();
      info.n_tape_ops = (int)c.read_tape.ops.size();
      if (!plan_only && !outs.empty()) {
        const Tensor *ref = get(G(tr.ops[c.read_ops[0]].out));
        bind_tape(c.read_tape, cx, rt);
        st = b200_launch_elemwise(&rt.tape, ins.data(), (int)ins.size(), outs.data(), (int)outs.size(),
                                  (int)ref->shape.size(), ref->shape.data(), nullptr);
      }
    } else if (c.kind == B200H_BLOCK_REDUCE) {
      const RelOp &red = tr.ops[c.core];
      std::vector<int32_t> out_ids = c.write_ops.empty() ? std::vector<int32_t>{red.out} : c.write_tape.outputs;
      for (auto id : out_ids)
        if ((st = materialise(G(id))) != B200_OK) return st;
      std::vector<int32_t> in_ids = c.read_ops.empty() ? std::vector<int32_t>{red.in[0]} : c.read_tape.inputs;
      if ((st = gather_descs(in_ids, 0, cx, ins)) != B200_OK) return st;
      if (!c.write_ops.empty() && (st = gather_descs(c.write_tape.inputs, 1, cx, wins)) != B200_OK) return st;
      if ((st = gather_descs(out_ids, 0, cx, outs)) != B200_OK) return st;
      info.n_inputs = (int)(ins.size() + wins.size());
      info.n_outputs = (int)outs.size();
      info.n_tape_ops = (int)(c.read_tape.ops.size() + c.write_tape.ops.size());
      if (!plan_only) {
        const Tensor *rin = get(G(red.in[0]));
        bind_tape(c.read_tape, cx, rt);
        bind_tape(c.write_tape, cx, wt);
        st = b200_launch_reduce(red.opcode, red.dim, (int)rin->shape.size(), rin->shape.data(),
                                c.read_ops.empty() ? nullptr : &rt.tape, ins.data(), (int)ins.size(),
                                c.write_ops.empty() ? nullptr : &wt.tape, wins.empty() ? nullptr : wins.data(),
                                (int)wins.size(), outs.data(), (int)outs.size(), nullptr);
      }
    } else if (c.kind == B200H_BLOCK_ROWNORM) {
      if ((st = materialise(G(c.row_out))) != B200_OK) return st;
      b200_tensor xin, yout;
      if ((st = desc_of(G(c.row_in), xin)) != B200_OK || (st = desc_of(G(c.row_out), yout)) != B200_OK) return st;
      info.n_inputs = 1 + (c.row_gamma >= 0) + (c.row_beta >= 0);
      info.n_outputs = 1;
      if (c.row_kind == 2) {
        b200_tensor g, bt;
        if ((st = desc_of(G(c.row_gamma), g)) != B200_OK) return st;
        if (c.row_beta >= 0 && (st = desc_of(G(c.row_beta), bt)) != B200_OK) return st;
        if (!plan_only) st = b200_launch_layer_norm(&xin, &g, c.row_beta >= 0 ? &bt : nullptr, cx.scalar[c.row_eps], &yout, nullptr);
      } else if (!plan_only) {
        st = b200_launch_softmax(&xin, &yout, c.row_kind == 1 ? 1 : 0, nullptr);
      }
    } else if (c.kind == B200H_BLOCK_MATMUL) {
      const RelOp &mm = tr.ops[c.core];
      const int64_t out_id = G(c.write_ops.empty() ? mm.out : c.write_tape.outputs[0]);
      if ((st = materialise(out_id)) != B200_OK) return st;
      b200_tensor a, b, cc;
      if ((st = desc_of(G(mm.in[0]), a)) != B200_OK || (st = desc_of(G(mm.in[1]), b)) != B200_OK ||
          (st = desc_of(out_id, cc)) != B200_OK)
        return st;
      if (!c.write_ops.empty() && (st = gather_descs(c.write_tape.inputs, 1, cx, wins)) != B200_OK) return st;
      info.n_inputs = 2 + (int)wins.size();
      info.n_outputs = 1;
      info.n_tape_ops = (int)c.write_tape.ops.size();
      if (!plan_only) {
        uint64_t wsb = 0;
        if ((st = b200_matmul_workspace_bytes(&a, &b, mm.precision, &wsb)) != B200_OK) return st;
        void *ws = nullptr;
        if (wsb && (st = b200_alloc(&ws, wsb, nullptr)) != B200_OK) return st;
        bind_tape(c.write_tape, cx, wt);
        st = b200_launch_matmul(&a, &b, &cc, mm.precision, c.write_ops.empty() ? nullptr : &wt.tape,
                                wins.empty() ? nullptr : wins.data(), (int)wins.size(), ws, wsb, nullptr);
        if (ws) b200_free(ws, nullptr);
      }
    } else {
      return fail(B200_ERR_INVALID, "unknown optimization kind %d", c.kind);
    }
    if (st != B200_OK) return st;
    info.launches = (int32_t)(b200_launch_count() - before) + (plan_only ? extra_launches : 0);
    log.push_back(info);
    for (auto id : consumed_ids) tensors.erase(id);  // R::free_handle for consumed ReadWrite handles
    queue.erase(queue.begin(), queue.begin() + c.consumed);
    return B200_OK;
  }

  // ---- the Explorer: offer the queue head to every fuser until all are closed, the best ready one wins
  std::shared_ptr<Optimization> explore() {
    const Op &head = queue[0];
    if (head.kind == Kind::Drop || is_view(head.kind) || head.kind == Kind::Gather || head.kind == Kind::Select) {
      auto o = std::make_shared<Optimization>();
      o->kind = head.kind == Kind::Drop ? (int)kBlockDrop : is_view(head.kind) ? (int)B200H_BLOCK_VIEW : (int)B200H_BLOCK_EAGER;
      o->consumed = 1;
      o->n_ops = head.kind == Kind::Drop ? 0 : 1;
      o->core = 0;
      Trace tr;
      Context cx;
      relativise(1, tr, cx);
      o->trace = tr;
      return o;
    }
    const size_t window = std::min(queue.size(), kExploreWindow);
    Trace tr;
    Context cx;
    relativise(window, tr, cx);
    // priority on ties: ReduceBroadcasted, Matmul, Reduce, ElementWise
    std::unique_ptr<Fuser> fusers[4] = {make_fuser(B200H_BLOCK_ROWNORM), make_fuser(B200H_BLOCK_MATMUL),
                                        make_fuser(B200H_BLOCK_REDUCE), make_fuser(B200H_BLOCK_ELEMWISE)};
    for (size_t i = 0; i < window; ++i) {
      bool open = false;
      for (auto &f : fusers) {
        f->offer(tr, (int)i);
        open = open || f->status == B200H_FUSER_OPEN;
      }
      if (!open) break;
    }
    Fuser *win = nullptr;
    for (auto &f : fusers)
      if (f->ready() && (!win || f->score() > win->score())) win = f.get();
    return win ? win->finish(tr) : nullptr;
  }

  int32_t drain() {
    if (queue.empty()) return B200_OK;
    std::string key;
    {
      Trace full;
      Context cx;
      relativise(queue.size(), full, cx);
      key = trace_key(full);
    }
    auto hit = plans.find(key);
    if (hit != plans.end()) {
      ++hits;
      bool stale = false;
      for (auto &o : hit->second) {
        int32_t st = run(*o, true);
        if (st == kMismatch) { stale = true; break; }
        if (st != B200_OK) return st;
      }
      if (!stale && queue.empty()) return B200_OK;
      plans.erase(key);  // the plan no longer fits what the fusers see: explore again and replace it
      key.clear();
    } else {
      ++misses;
    }
    std::vector<std::shared_ptr<Optimization>> plan;
    while (!queue.empty()) {
      auto o = explore();
      B200_REQUIRE(o, B200_ERR_UNSUPPORTED, "no fuser accepts operation kind %d", (int)queue[0].kind);
      int32_t st = run(*o, false);
      B200_REQUIRE(st != kMismatch, B200_ERR_INVALID, "internal: a fresh optimization does not match its own operations");
      if (st != B200_OK) return st;
      plan.push_back(o);
    }
    if (!key.empty()) plans[key] = std::move(plan);
    return B200_OK;
  }
};

static int32_t bshape(const std::vector<int64_t> &a, const std::vector<int64_t> &b, std::vector<int64_t> &out) {
  B200_REQUIRE(a.size() == b.size(), B200_ERR_SHAPE, "rank mismatch %zu vs %zu", a.size(), b.size());
  out.resize(a.size());
  for (size_t i = 0; i < a.size(); ++i) {
    B200_REQUIRE(a[i] == b[i] || a[i] == 1 || b[i] == 1, B200_ERR_SHAPE, "dim %zu: %lld vs %lld not broadcastable", i,
                 (long long)a[i], (long long)b[i]);
    out[i] = std::max(a[i], b[i]);
  }
  return B200_OK;
}

// the exported OperationFuser handle: a fuser bound to the relative form of one stream's pending queue
struct FuserHandle {
  std::shared_ptr<Trace> trace;
  std::unique_ptr<Fuser> fuser;
};

}  // namespace b200h

using namespace b200h;
#define S(s) (reinterpret_cast<Stream *>(s))
#define NEED(s, id, var)                                                               \
  Tensor *var = S(s)->get(id);                                                         \
  if (!var) { b200::fail(B200_ERR_INVALID, "unknown tensor id %lld", (long long)(id)); return -1; }

extern "C" {

int32_t b200h_stream_create(b200h_stream *out, int32_t plan_only) {
  B200_REQUIRE(out, B200_ERR_INVALID, "out is null");
  if (!plan_only) {
    int32_t st = b200_init(0);
    if (st != B200_OK) return st;
  }
  Stream *s = new Stream();
  s->plan_only = plan_only != 0;
  *out = s;
  return B200_OK;
}

int32_t b200h_stream_destroy(b200h_stream s) {
  if (s) {
    if (!S(s)->plan_only) b200_stream_sync(nullptr);
    delete S(s);
  }
  return B200_OK;
}

b200h_id b200h_from_host(b200h_stream s, const void *data, int32_t dtype, int32_t rank, const int64_t *shape) {
  if (!s || rank < 1 || rank > B200_MAX_RANK || !shape || b200::dtype_size(dtype) == 0) {
    b200::fail(B200_ERR_INVALID, "bad arguments to b200h_from_host");
    return -1;
  }
  std::vector<int64_t> sh(shape, shape + rank);
  const int64_t id = S(s)->add_pending(sh, dtype);
  if (S(s)->materialise(id) != B200_OK) return -1;
  if (!S(s)->plan_only && data) {
    Tensor *t = S(s)->get(id);
    const uint64_t bytes = (uint64_t)numel(sh) * b200::dtype_size(dtype);
    if (b200_memcpy_h2d(t->buf->ptr, data, bytes, nullptr) != B200_OK) return -1;
    if (b200_stream_sync(nullptr) != B200_OK) return -1;  // `data` may be a temporary
  }
  return id;
}

b200h_id b200h_from_device(b200h_stream s, void *ptr, int32_t dtype, int32_t rank, const int64_t *shape, const int64_t *strides) {
  if (!s || rank < 1 || rank > B200_MAX_RANK || !shape || b200::dtype_size(dtype) == 0 || (!ptr && !S(s)->plan_only)) {
    b200::fail(B200_ERR_INVALID, "bad arguments to b200h_from_device");
    return -1;
  }
  Tensor t;
  t.shape.assign(shape, shape + rank);
  t.strides = strides ? std::vector<int64_t>(strides, strides + rank) : contiguous(t.shape);
  t.dtype = dtype;
  t.buf = std::make_shared<Buffer>();
  t.buf->ptr = ptr;
  t.buf->owned = false;
  t.buf->fake = S(s)->plan_only;
  const int64_t id = S(s)->next_id++;
  S(s)->tensors[id] = std::move(t);
  return id;
}

int32_t b200h_device_tensor(b200h_stream s, b200h_id id, b200_tensor *out) {
  B200_REQUIRE(s && out, B200_ERR_INVALID, "null argument");
  int32_t st = S(s)->drain();
  if (st != B200_OK) return st;
  return S(s)->desc_of(id, *out);
}

int32_t b200h_sync(b200h_stream s) {
  int32_t st = S(s)->drain();
  if (st != B200_OK) return st;
  return S(s)->plan_only ? B200_OK : b200_stream_sync(nullptr);
}

int32_t b200h_flush(b200h_stream s) { return S(s)->drain(); }

int32_t b200h_shape(b200h_stream s, b200h_id id, int32_t *dtype, int32_t *rank, int64_t *shape) {
  Tensor *t = S(s)->get(id);
  B200_REQUIRE(t, B200_ERR_INVALID, "unknown tensor id %lld", (long long)id);
  if (dtype) *dtype = t->dtype;
  if (rank) *rank = (int32_t)t->shape.size();
  if (shape) for (size_t i = 0; i < t->shape.size(); ++i) shape[i] = t->shape[i];
  return B200_OK;
}

int32_t b200h_read(b200h_stream s, b200h_id id, void *dst, uint64_t dst_bytes) {
  int32_t st = S(s)->drain();
  if (st != B200_OK) return st;
  Tensor *t = S(s)->get(id);
  B200_REQUIRE(t && t->buf, B200_ERR_INVALID, "tensor %lld was dropped or never computed", (long long)id);
  B200_REQUIRE(!S(s)->plan_only, B200_ERR_UNSUPPORTED, "plan-only streams hold no data");
  const uint64_t bytes = (uint64_t)numel(t->shape) * b200::dtype_size(t->dtype);
  B200_REQUIRE(dst && dst_bytes >= bytes, B200_ERR_INVALID, "destination too small");
  b200_tensor src;
  if ((st = S(s)->desc_of(id, src)) != B200_OK) return st;
  if (b200::is_contiguous(src)) {
    st = b200_memcpy_d2h(dst, src.ptr, bytes, nullptr);
  } else {
    void *tmp = nullptr;
    if ((st = b200_alloc(&tmp, bytes, nullptr)) != B200_OK) return st;
    b200_tensor d = src;
    d.ptr = tmp;
    auto cs = contiguous(t->shape);
    for (size_t i = 0; i < cs.size(); ++i) d.strides[i] = cs[i];
    st = b200_launch_copy(&src, &d, nullptr);
    if (st == B200_OK) st = b200_memcpy_d2h(dst, tmp, bytes, nullptr);
    b200_free(tmp, nullptr);
  }
  if (st != B200_OK) return st;
  return b200_stream_sync(nullptr);
}

static b200h_id push_elemwise(b200h_stream s, Kind kind, int opcode, b200h_id a, b200h_id b, b200h_id c, double scalar) {
  NEED(s, a, ta);
  std::vector<int64_t> shape = ta->shape;
  int32_t dtype = ta->dtype;
  for (b200h_id other : {b, c}) {
    if (other < 0) continue;
    NEED(s, other, to);
    std::vector<int64_t> merged;
    if (bshape(shape, to->shape, merged) != B200_OK) return -1;
    shape = merged;
  }
  if (kind != Kind::MaskFill && kind != Kind::MaskWhere) {
    if (opcode < 0 || opcode >= B200_OP_COUNT) { b200::fail(B200_ERR_INVALID, "bad opcode %d", opcode); return -1; }
    if (is_cmp(opcode)) dtype = B200_BOOL;
    else if (opcode == B200_OP_F2I) dtype = B200_I32;
    else if (opcode == B200_OP_I2F || opcode == B200_OP_B2F) dtype = B200_F32;
  }
  Op o;
  o.kind = kind;
  o.opcode = opcode;
  o.in[0] = a; o.in[1] = b; o.in[2] = c;
  o.scalar = scalar;
  o.out = S(s)->add_pending(shape, dtype);
  S(s)->queue.push_back(o);
  return o.out;
}

b200h_id b200h_binary(b200h_stream s, int32_t opcode, b200h_id lhs, b200h_id rhs) {
  return push_elemwise(s, Kind::Binary, opcode, lhs, rhs, -1, 0);
}
b200h_id b200h_scalar(b200h_stream s, int32_t opcode, b200h_id lhs, double scalar) {
  return push_elemwise(s, Kind::Scalar, opcode, lhs, -1, -1, scalar);
}
b200h_id b200h_unary(b200h_stream s, int32_t opcode, b200h_id x) {
  return push_elemwise(s, Kind::Unary, opcode, x, -1, -1, 0);
}
b200h_id b200h_mask_fill(b200h_stream s, b200h_id x, b200h_id mask, double value) {
  return push_elemwise(s, Kind::MaskFill, B200_OP_SELECT, x, mask, -1, value);
}
b200h_id b200h_mask_where(b200h_stream s, b200h_id x, b200h_id mask, b200h_id source) {
  return push_elemwise(s, Kind::MaskWhere, B200_OP_SELECT, x, mask, source, 0);
}

b200h_id b200h_reduce_dim(b200h_stream s, int32_t kind, b200h_id x, int32_t dim) {
  NEED(s, x, t);
  if (dim < 0) dim += (int32_t)t->shape.size();
  if (dim < 0 || dim >= (int32_t)t->shape.size()) {
    b200::fail(B200_ERR_SHAPE, "dim %d out of range for rank %zu", dim, t->shape.size());
    return -1;
  }
  std::vector<int64_t> shape = t->shape;
  shape[dim] = 1;
  int32_t dtype = t->dtype;
  if (kind == B200_RED_ARGMAX || kind == B200_RED_ARGMIN) dtype = B200_I32;  // CUDA default IntElem
  if (kind == B200_RED_ANY || kind == B200_RED_ALL) dtype = B200_BOOL;
  Op o;
  o.kind = Kind::ReduceDim;
  o.opcode = kind;
  o.in[0] = x;
  o.dim = dim;
  o.out = S(s)->add_pending(shape, dtype);
  S(s)->queue.push_back(o);
  return o.out;
}

b200h_id b200h_matmul(b200h_stream s, b200h_id lhs, b200h_id rhs, int32_t precision) {
  NEED(s, lhs, a);
  NEED(s, rhs, b);
  const size_t r = a->shape.size();
  if (r != b->shape.size() || r < 2 || a->shape[r - 1] != b->shape[r - 2]) {
    b200::fail(B200_ERR_SHAPE, "matmul shapes are incompatible");
    return -1;
  }
  std::vector<int64_t> shape(r);
  for (size_t d = 0; d + 2 < r; ++d) {
    if (a->shape[d] != b->shape[d] && a->shape[d] != 1 && b->shape[d] != 1) {
      b200::fail(B200_ERR_SHAPE, "matmul batch dim %zu not broadcastable", d);
      return -1;
    }
    shape[d] = std::max(a->shape[d], b->shape[d]);
  }
  shape[r - 2] = a->shape[r - 2];
  shape[r - 1] = b->shape[r - 1];
  Op o;
  o.kind = Kind::Matmul;
  o.in[0] = lhs; o.in[1] = rhs;
  o.precision = precision;
  o.out = S(s)->add_pending(shape, B200_F32);
  S(s)->queue.push_back(o);
  return o.out;
}

// A view of a tensor that already has storage is resolved at once (no queue entry, nothing to launch); a view of a
// pending tensor — or a reshape that has to copy — is queued and runs as a lone block.
static b200h_id push_view(b200h_stream s, Op o, const std::vector<int64_t> &out_shape, bool needs_copy) {
  Tensor *src = S(s)->get(o.in[0]);
  Tensor t;
  t.shape = out_shape;
  t.strides = contiguous(out_shape);
  t.dtype = src->dtype;
  o.out = S(s)->next_id++;
  const bool now = src->buf && !needs_copy;
  S(s)->tensors[o.out] = std::move(t);
  if (now) {
    if (S(s)->apply_view(o, nullptr) != B200_OK) return -1;
  } else {
    S(s)->queue.push_back(o);
  }
  return o.out;
}

b200h_id b200h_swap_dims(b200h_stream s, b200h_id x, int32_t d0, int32_t d1) {
  // metadata-only view (crates/burn-cubecl/src/ops/base.rs:137-139)
  NEED(s, x, t);
  const int r = (int)t->shape.size();
  if (d0 < 0) d0 += r;
  if (d1 < 0) d1 += r;
  if (d0 < 0 || d1 < 0 || d0 >= r || d1 >= r) { b200::fail(B200_ERR_SHAPE, "swap_dims out of range"); return -1; }
  Op o;
  o.kind = Kind::SwapDims;
  o.in[0] = x;
  o.dim = d0;
  o.dim2 = d1;
  std::vector<int64_t> shape = t->shape;
  std::swap(shape[d0], shape[d1]);
  return push_view(s, o, shape, false);
}

b200h_id b200h_reshape(b200h_stream s, b200h_id x, int32_t rank, const int64_t *shape) {
  NEED(s, x, t);
  if (rank < 1 || rank > B200_MAX_RANK || !shape) { b200::fail(B200_ERR_INVALID, "bad reshape arguments"); return -1; }
  std::vector<int64_t> sh(shape, shape + rank);
  if (numel(sh) != numel(t->shape)) {
    b200::fail(B200_ERR_SHAPE, "reshape changes the element count (%lld -> %lld)", (long long)numel(t->shape), (long long)numel(sh));
    return -1;
  }
  Op o;
  o.kind = Kind::Reshape;
  o.in[0] = x;
  return push_view(s, o, sh, t->buf && !is_contig(t->shape, t->strides));
}

b200h_id b200h_expand(b200h_stream s, b200h_id x, int32_t rank, const int64_t *shape) {
  NEED(s, x, t);
  const int r = (int)t->shape.size();
  if (rank < r || rank > B200_MAX_RANK || !shape) { b200::fail(B200_ERR_INVALID, "bad expand arguments"); return -1; }
  std::vector<int64_t> sh(shape, shape + rank);
  for (int d = 0; d < r; ++d) {
    const int64_t have = t->shape[d], want = sh[rank - r + d];
    if (have != want && have != 1) { b200::fail(B200_ERR_SHAPE, "expand: dim %d is %lld, cannot become %lld", d, (long long)have, (long long)want); return -1; }
  }
  Op o;
  o.kind = Kind::Expand;
  o.in[0] = x;
  return push_view(s, o, sh, false);
}

b200h_id b200h_slice(b200h_stream s, b200h_id x, const int64_t *starts, const int64_t *ends) {
  NEED(s, x, t);
  if (!starts || !ends) { b200::fail(B200_ERR_INVALID, "bad slice arguments"); return -1; }
  const size_t r = t->shape.size();
  Op o;
  o.kind = Kind::Slice;
  o.in[0] = x;
  std::vector<int64_t> sh(r);
  for (size_t d = 0; d < r; ++d) {
    if (starts[d] < 0 || ends[d] < starts[d] || ends[d] > t->shape[d]) {
      b200::fail(B200_ERR_SHAPE, "slice: dim %zu range %lld..%lld outside 0..%lld", d, (long long)starts[d], (long long)ends[d], (long long)t->shape[d]);
      return -1;
    }
    sh[d] = ends[d] - starts[d];
  }
  o.range.assign(starts, starts + r);
  o.range.insert(o.range.end(), ends, ends + r);
  return push_view(s, o, sh, false);
}

static b200h_id push_indexed(b200h_stream s, Kind kind, int32_t dim, b200h_id x, b200h_id indices) {
  NEED(s, x, t);
  NEED(s, indices, ix);
  const int r = (int)t->shape.size();
  if (dim < 0) dim += r;
  if (dim < 0 || dim >= r) { b200::fail(B200_ERR_SHAPE, "dim %d out of range for rank %d", dim, r); return -1; }
  if (ix->dtype != B200_I32 && ix->dtype != B200_I64) { b200::fail(B200_ERR_INVALID, "indices must be i32 or i64"); return -1; }
  std::vector<int64_t> shape;
  if (kind == Kind::Gather) {
    if ((int)ix->shape.size() != r) { b200::fail(B200_ERR_SHAPE, "gather: indices rank %zu != tensor rank %d", ix->shape.size(), r); return -1; }
    shape = ix->shape;
  } else {
    if (ix->shape.size() != 1) { b200::fail(B200_ERR_SHAPE, "select: indices must be 1-D"); return -1; }
    shape = t->shape;
    shape[dim] = ix->shape[0];
  }
  Op o;
  o.kind = kind;
  o.in[0] = x;
  o.in[1] = indices;
  o.dim = dim;
  o.out = S(s)->add_pending(shape, t->dtype);
  S(s)->queue.push_back(o);
  return o.out;
}
b200h_id b200h_gather(b200h_stream s, int32_t dim, b200h_id x, b200h_id indices) { return push_indexed(s, Kind::Gather, dim, x, indices); }
b200h_id b200h_select(b200h_stream s, int32_t dim, b200h_id x, b200h_id indices) { return push_indexed(s, Kind::Select, dim, x, indices); }

int32_t b200h_drop(b200h_stream s, b200h_id id) {
  B200_REQUIRE(S(s)->get(id), B200_ERR_INVALID, "unknown tensor id %lld", (long long)id);
  if (S(s)->queue.empty()) {  // nothing pending can still read it: release the handle now
    S(s)->tensors.erase(id);
    return B200_OK;
  }
  Op o;
  o.kind = Kind::Drop;
  o.in[0] = id;
  S(s)->queue.push_back(o);
  return B200_OK;
}

// ---- plan cache
int32_t b200h_cache_stats_get(b200h_stream s, b200h_cache_stats *out) {
  B200_REQUIRE(s && out, B200_ERR_INVALID, "null argument");
  out->hits = S(s)->hits;
  out->misses = S(s)->misses;
  out->plans = S(s)->plans.size();
  out->inplace_aliases = g_inplace_aliases.load();
  return B200_OK;
}
int32_t b200h_cache_clear(b200h_stream s) {
  S(s)->plans.clear();
  S(s)->hits = S(s)->misses = 0;
  return B200_OK;
}

// ---- OperationFuser / Optimization handles
#define FH(f) (reinterpret_cast<FuserHandle *>(f))
#define OH(o) (reinterpret_cast<std::shared_ptr<Optimization> *>(o))

int32_t b200h_fuser_create(b200h_stream s, int32_t block_kind, b200h_fuser *out) {
  B200_REQUIRE(s && out, B200_ERR_INVALID, "null argument");
  auto f = make_fuser(block_kind);
  B200_REQUIRE(f, B200_ERR_INVALID, "no fuser of kind %d", block_kind);
  auto *h = new FuserHandle();
  h->fuser = std::move(f);
  h->trace = std::make_shared<Trace>();
  Context cx;
  S(s)->relativise(std::min(S(s)->queue.size(), kExploreWindow), *h->trace, cx);
  *out = h;
  return B200_OK;
}
int32_t b200h_fuser_destroy(b200h_fuser f) {
  delete FH(f);
  return B200_OK;
}
int32_t b200h_fuser_fuse_next(b200h_fuser f) {
  B200_REQUIRE(f, B200_ERR_INVALID, "null fuser");
  B200_REQUIRE(FH(f)->fuser->fed < (int)FH(f)->trace->ops.size(), B200_ERR_INVALID, "no more queued operations to offer");
  FH(f)->fuser->offer(*FH(f)->trace, FH(f)->fuser->fed);
  return B200_OK;
}
int32_t b200h_fuser_status_get(b200h_fuser f) { return f ? FH(f)->fuser->status : (int32_t)B200H_FUSER_CLOSED; }
int32_t b200h_fuser_properties(b200h_fuser f, uint64_t *score, int32_t *ready) {
  B200_REQUIRE(f, B200_ERR_INVALID, "null fuser");
  if (score) *score = FH(f)->fuser->score();
  if (ready) *ready = FH(f)->fuser->ready() ? 1 : 0;
  return B200_OK;
}
int32_t b200h_fuser_len(b200h_fuser f) { return f ? FH(f)->fuser->len() : 0; }
int32_t b200h_fuser_reset(b200h_fuser f) {
  B200_REQUIRE(f, B200_ERR_INVALID, "null fuser");
  FH(f)->fuser->reset();
  return B200_OK;
}
int32_t b200h_fuser_clone(b200h_fuser f, b200h_fuser *out) {
  B200_REQUIRE(f && out, B200_ERR_INVALID, "null argument");
  auto *h = new FuserHandle();
  h->trace = FH(f)->trace;
  h->fuser = FH(f)->fuser->clone();
  *out = h;
  return B200_OK;
}
int32_t b200h_fuser_finish(b200h_fuser f, b200h_optimization *out) {
  B200_REQUIRE(f && out, B200_ERR_INVALID, "null argument");
  auto o = FH(f)->fuser->finish(*FH(f)->trace);
  B200_REQUIRE(o, B200_ERR_INVALID, "the fuser is not ready: nothing to finish");
  *out = new std::shared_ptr<Optimization>(std::move(o));
  return B200_OK;
}
int32_t b200h_optimization_destroy(b200h_optimization o) {
  delete OH(o);
  return B200_OK;
}
int32_t b200h_optimization_len(b200h_optimization o) { return o ? (*OH(o))->n_ops : 0; }
const char *b200h_optimization_name(b200h_optimization o) { return o ? (*OH(o))->name() : ""; }
int32_t b200h_optimization_execute(b200h_optimization o, b200h_stream s) {
  B200_REQUIRE(o && s, B200_ERR_INVALID, "null argument");
  int32_t st = S(s)->run(**OH(o), false);
  B200_REQUIRE(st != kMismatch, B200_ERR_INVALID,
               "the stream's pending operations are not the ones this optimization was built from");
  return st;
}
int32_t b200h_optimization_to_state(b200h_optimization o, void *buf, uint64_t cap, uint64_t *len) {
  B200_REQUIRE(o && len, B200_ERR_INVALID, "null argument");
  const std::string st = (*OH(o))->to_state();
  *len = st.size();
  if (buf && cap >= st.size()) memcpy(buf, st.data(), st.size());
  else if (buf) return fail(B200_ERR_INVALID, "state needs %zu bytes, buffer has %llu", st.size(), (unsigned long long)cap);
  return B200_OK;
}
int32_t b200h_optimization_from_state(const void *buf, uint64_t len, b200h_optimization *out) {
  B200_REQUIRE(buf && out, B200_ERR_INVALID, "null argument");
  auto o = Optimization::from_state(buf, len);
  B200_REQUIRE(o, B200_ERR_INVALID, "not a valid optimization state");
  *out = new std::shared_ptr<Optimization>(std::move(o));
  return B200_OK;
}

int32_t b200h_block_count(b200h_stream s) { return (int32_t)S(s)->log.size(); }
int32_t b200h_block_get(b200h_stream s, int32_t index, b200h_block_info *out) {
  B200_REQUIRE(out && index >= 0 && index < (int32_t)S(s)->log.size(), B200_ERR_INVALID, "bad block index");
  *out = S(s)->log[index];
  return B200_OK;
}
int32_t b200h_block_clear(b200h_stream s) {
  S(s)->log.clear();
  return B200_OK;
}

}  // extern "C"
