// Host-side fusion layer: a lazy operation stream, the three OperationFuser state machines
// (ElementWise, Matmul, Reduce) and Optimization::execute on top of the kernel ABI.
//
// Mirrors, for the hot path only, what burn-fusion + burn-cubecl-fusion do in the reference:
//   OperationIr stream     crates/burn-ir/src/operation.rs:113-142
//   OperationFuser         crates/burn-fusion/src/backend.rs:187-206  (fuse / status / len / finish)
//   acceptance rules       crates/burn-cubecl-fusion/src/engine/fuser.rs:76-190,292-710 (same output
//                          shape, <= 64 ops, bounded bindings, Drop absorbed so intermediates stay in
//                          registers), optim/reduce/fuser.rs:103-160,221-300 (read block → one *Dim
//                          reduce → write block), optim/matmul/fuser.rs:69-151 (matmul + epilogue)
//   block choice           the candidate absorbing the most operations wins (scoring.rs:56-76 rewards
//                          saved launches and saved global IO, both monotone in the op count here)
//   Optimization::execute  crates/burn-fusion/src/backend.rs:226-234 — resolve ids to handles, allocate
//                          outputs, ONE launch per block, register outputs, free dropped handles.
// The stream/plan-cache/beam-search machinery of burn-fusion itself is out of scope.
#include <algorithm>
#include <cstring>
#include <map>
#include <memory>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../csrc/common.cuh"
#include "burn_b200_host.h"

namespace b200h {

using b200::fail;

struct Buffer {
  void *ptr = nullptr;
  bool fake = false;
  ~Buffer() {
    if (ptr && !fake) b200_free(ptr, nullptr);
  }
};

struct Tensor {
  std::vector<int64_t> shape, strides;
  int32_t dtype = B200_F32;
  std::shared_ptr<Buffer> buf;  // null while the producing op is still queued
  int64_t offset = 0;           // elements
};

enum class Kind { Binary, Scalar, Unary, MaskFill, MaskWhere, ReduceDim, Matmul, Drop };

struct Op {
  Kind kind;
  int opcode = 0;       // b200_opcode / b200_reduce_kind
  int64_t in[3] = {-1, -1, -1};
  int64_t out = -1;
  double scalar = 0;
  int dim = 0;
  int precision = 0;
};

static bool is_cmp(int op) {
  return (op >= B200_OP_EQ_F && op <= B200_OP_ISINF_F) || (op >= B200_OP_EQ_I && op <= B200_OP_GE_I) ||
         (op >= B200_OP_AND_B && op <= B200_OP_NOT_B) || op == B200_OP_F2B || op == B200_OP_I2B;
}
static bool is_int_op(int op) { return op >= B200_OP_ADD_I && op <= B200_OP_GE_I; }
static bool is_elemwise(Kind k) {
  return k == Kind::Binary || k == Kind::Scalar || k == Kind::Unary || k == Kind::MaskFill || k == Kind::MaskWhere;
}

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

// ---------------------------------------------------------------- tape assembly
struct TapeBuild {
  std::vector<b200_tape_op> ops;
  std::vector<uint32_t> scalars;
  std::vector<int64_t> inputs;   // global tensor ids, in INPUT(k) order
  std::vector<int64_t> outputs;  // tensor ids written, in out-index order
  bool ok = false;
};

static uint32_t f32_bits(double v) {
  float f = (float)v;
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}

// Builds the public tape for a run of elementwise ops.  `virtual_in0` (>= 0) is a tensor id
// that must map to INPUT(0) without being a real global input (reduced value / accumulator).
static TapeBuild build_tape(const std::vector<const Op *> &ops, const std::unordered_set<int64_t> &dropped,
                            int64_t virtual_in0, bool virtual_needs_output, int max_outputs) {
  TapeBuild tb;
  std::unordered_map<int64_t, int> producer;  // tensor id -> op index in block
  for (size_t i = 0; i < ops.size(); ++i) producer[ops[i]->out] = (int)i;
  // last use of each local value
  std::vector<int> last_use(ops.size(), -1);
  std::vector<bool> far_use(ops.size(), false);
  for (size_t i = 0; i < ops.size(); ++i)
    for (int k = 0; k < 3; ++k) {
      auto it = producer.find(ops[i]->in[k]);
      if (it != producer.end() && it->second < (int)i) {
        last_use[it->second] = (int)i;
        if ((int)i > it->second + 1) far_use[it->second] = true;
      }
    }
  if (virtual_in0 >= 0) tb.inputs.push_back(virtual_in0);
  auto input_index = [&](int64_t id) -> int {
    for (size_t k = 0; k < tb.inputs.size(); ++k)
      if (tb.inputs[k] == id) return (int)k;
    tb.inputs.push_back(id);
    return (int)tb.inputs.size() - 1;
  };
  auto scalar_index = [&](uint32_t bits) -> int {
    for (size_t k = 0; k < tb.scalars.size(); ++k)
      if (tb.scalars[k] == bits) return (int)k;
    tb.scalars.push_back(bits);
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
    const Op &o = *ops[i];
    auto arg = [&](int64_t id) -> int {
      auto it = producer.find(id);
      if (it != producer.end() && it->second < (int)i) {
        const int j = it->second;
        if (j == (int)i - 1 && !(virtual_needs_output && false)) return B200_ARG_ACC;
        return B200_ARG_TEMP(temp_of[j]);
      }
      return B200_ARG_INPUT(input_index(id));
    };
    b200_tape_op t;
    memset(&t, 0, sizeof(t));
    t.dst_temp = B200_DST_NONE;
    t.dst_out = B200_DST_NONE;
    switch (o.kind) {
      case Kind::Binary:
        t.op = (uint8_t)o.opcode; t.a = (uint8_t)arg(o.in[0]); t.b = (uint8_t)arg(o.in[1]);
        break;
      case Kind::Scalar: {
        t.op = (uint8_t)o.opcode; t.a = (uint8_t)arg(o.in[0]);
        const uint32_t bits = is_int_op(o.opcode) ? (uint32_t)(int32_t)o.scalar : f32_bits(o.scalar);
        t.b = (uint8_t)B200_ARG_SCALAR(scalar_index(bits));
        break;
      }
      case Kind::Unary:
        t.op = (uint8_t)o.opcode; t.a = (uint8_t)arg(o.in[0]);
        break;
      case Kind::MaskFill:  // ConditionalAssign: cond ? scalar : x
        t.op = B200_OP_SELECT; t.a = (uint8_t)arg(o.in[0]);
        t.b = (uint8_t)B200_ARG_SCALAR(scalar_index(f32_bits(o.scalar))); t.c = (uint8_t)arg(o.in[1]);
        break;
      case Kind::MaskWhere:
        t.op = B200_OP_SELECT; t.a = (uint8_t)arg(o.in[0]); t.b = (uint8_t)arg(o.in[2]); t.c = (uint8_t)arg(o.in[1]);
        break;
      default: return tb;
    }
    // the first op of a plain block must not read ACC; a value used later than the next op needs a temp
    if (far_use[i] || (last_use[i] == (int)i + 1 && false)) {
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

// ---------------------------------------------------------------- candidates (fusers)
struct Candidate {
  int kind = B200H_BLOCK_EAGER;
  int consumed = 0;  // queue entries (ops + drops)
  int n_ops = 0;     // IR ops excluding drops
  std::vector<const Op *> read_ops, write_ops;
  const Op *core = nullptr;  // the reduce / matmul op
  int64_t row_in = -1, row_out = -1;   // ROWNORM: the chain's input and output tensors
  int row_kind = 0;                    // ROWNORM: 0 softmax, 1 log_softmax, 2 layer_norm
  int64_t row_gamma = -1, row_beta = -1;
  double row_eps = 0;
  std::unordered_set<int64_t> dropped;
  TapeBuild read_tape, write_tape;
};

struct Stream {
  bool plan_only = false;
  int64_t next_id = 1;
  std::unordered_map<int64_t, Tensor> tensors;
  std::vector<Op> queue;
  std::vector<b200h_block_info> log;

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

  // ---- ElementWise fuser: longest prefix of same-shape elementwise ops (drops absorbed)
  Candidate scan_elemwise(size_t start_shape_from = 0) {
    Candidate best, cur;
    cur.kind = B200H_BLOCK_ELEMWISE;
    std::vector<int64_t> shape;
    for (size_t q = 0; q < queue.size(); ++q) {
      const Op &o = queue[q];
      if (o.kind == Kind::Drop) {
        if (cur.read_ops.empty()) break;  // leading drops are applied by the drain loop
        cur.dropped.insert(o.in[0]);
        cur.consumed = (int)q + 1;
        TapeBuild tb = build_tape(cur.read_ops, cur.dropped, -1, false, B200_MAX_TAPE_OUTPUTS);
        if (tb.ok) { best = cur; best.read_tape = tb; }
        continue;
      }
      if (!is_elemwise(o.kind)) break;
      const Tensor *out = get(o.out);
      if (cur.read_ops.empty()) shape = out->shape;
      else if (out->shape != shape) break;  // output_is_compatible: same shape only
      if ((int)cur.read_ops.size() >= B200_MAX_TAPE_OPS) break;
      cur.read_ops.push_back(&o);
      cur.n_ops = (int)cur.read_ops.size();
      cur.consumed = (int)q + 1;
      TapeBuild tb = build_tape(cur.read_ops, cur.dropped, -1, false, B200_MAX_TAPE_OUTPUTS);
      if (!tb.ok) break;
      best = cur;
      best.read_tape = tb;
    }
    (void)start_shape_from;
    return best;
  }

  // ---- ReduceBroadcasted fuser: the softmax / log_softmax chains the default ActivationOps record
  //   softmax     = max_dim, sub, exp, sum_dim, div          (activation.rs:250-256)
  //   log_softmax = max_dim, sub, exp, sum_dim, log, sub     (activation.rs:271-276)
  // along the last axis, every intermediate dropped inside the window → one row-resident kernel
  // (crates/burn-cubecl-fusion/src/optim/reduce_broadcasted/ fuses the same shape of work).
  Candidate scan_rownorm() {
    Candidate none, c;
    c.kind = B200H_BLOCK_ROWNORM;
    std::vector<const Op *> ops;
    std::vector<size_t> pos;
    for (size_t q = 0; q < queue.size() && ops.size() < 9; ++q) {
      if (queue[q].kind == Kind::Drop) { if (ops.empty()) return none; continue; }
      ops.push_back(&queue[q]);
      pos.push_back(q);
    }
    if (ops.size() < 5) return none;
    const Op &o0 = *ops[0];
    const Tensor *x = get(o0.in[0]);
    if (!x || x->dtype != B200_F32 || x->shape.empty() || x->strides != contiguous(x->shape)) return none;
    const int last = (int)x->shape.size() - 1;
    auto is_bin = [](const Op &o, int opc, int64_t a, int64_t b) { return o.kind == Kind::Binary && o.opcode == opc && o.in[0] == a && o.in[1] == b; };
    auto is_un = [](const Op &o, int opc, int64_t a) { return o.kind == Kind::Unary && o.opcode == opc && o.in[0] == a; };
    auto is_red = [&](const Op &o, int kind, int64_t a) { return o.kind == Kind::ReduceDim && o.opcode == kind && o.dim == last && o.in[0] == a; };
    std::vector<int64_t> inter;
    size_t n_ops = 0;
    const int64_t xid = o0.in[0];
    if (is_red(o0, B200_RED_MAX, xid) && is_bin(*ops[1], B200_OP_SUB_F, xid, o0.out) && is_un(*ops[2], B200_OP_EXP_F, ops[1]->out) &&
        is_red(*ops[3], B200_RED_SUM, ops[2]->out)) {
      inter = {o0.out, ops[1]->out, ops[2]->out, ops[3]->out};
      if (is_bin(*ops[4], B200_OP_DIV_F, ops[2]->out, ops[3]->out)) {
        n_ops = 5;
        c.row_kind = 0;
        c.row_out = ops[4]->out;
      } else if (ops.size() >= 6 && is_un(*ops[4], B200_OP_LOG_F, ops[3]->out) &&
                 is_bin(*ops[5], B200_OP_SUB_F, ops[1]->out, ops[4]->out)) {
        n_ops = 6;
        c.row_kind = 1;
        c.row_out = ops[5]->out;
        inter.push_back(ops[4]->out);
      } else {
        return none;
      }
    } else if (ops.size() >= 8 && is_red(o0, B200_RED_MEAN, xid) && is_bin(*ops[1], B200_OP_SUB_F, xid, o0.out) &&
               is_bin(*ops[2], B200_OP_MUL_F, ops[1]->out, ops[1]->out) && is_red(*ops[3], B200_RED_MEAN, ops[2]->out) &&
               ops[4]->kind == Kind::Scalar && ops[4]->opcode == B200_OP_ADD_F && ops[4]->in[0] == ops[3]->out &&
               is_un(*ops[5], B200_OP_SQRT_F, ops[4]->out) && is_bin(*ops[6], B200_OP_DIV_F, ops[1]->out, ops[5]->out) &&
               ops[7]->kind == Kind::Binary && ops[7]->opcode == B200_OP_MUL_F && ops[7]->in[0] == ops[6]->out) {
      // layer_norm = mean_dim, sub, mul, mean_dim, add_scalar, sqrt, div, mul(gamma) [, add(beta)]  (modules/base.rs:846-877)
      auto is_vec = [&](int64_t id) {
        const Tensor *t = get(id);
        return t && t->dtype == B200_F32 && numel(t->shape) == x->shape[last] && !t->shape.empty() && t->shape.back() == x->shape[last] &&
               t->strides == contiguous(t->shape);
      };
      if (!is_vec(ops[7]->in[1])) return none;
      inter = {o0.out, ops[1]->out, ops[2]->out, ops[3]->out, ops[4]->out, ops[5]->out, ops[6]->out};
      c.row_kind = 2;
      c.row_gamma = ops[7]->in[1];
      c.row_eps = ops[4]->scalar;
      n_ops = 8;
      c.row_out = ops[7]->out;
      if (ops.size() >= 9 && ops[8]->kind == Kind::Binary && ops[8]->opcode == B200_OP_ADD_F && ops[8]->in[0] == ops[7]->out &&
          is_vec(ops[8]->in[1])) {
        inter.push_back(ops[7]->out);
        c.row_beta = ops[8]->in[1];
        c.row_out = ops[8]->out;
        n_ops = 9;
      }
    } else {
      return none;
    }
    // drops seen up to the last op of the chain, plus the drops of intermediates that directly follow it
    std::unordered_set<int64_t> drops;
    size_t end = pos[n_ops - 1] + 1;
    for (size_t i = 0; i < end; ++i)
      if (queue[i].kind == Kind::Drop) drops.insert(queue[i].in[0]);
    while (end < queue.size() && queue[end].kind == Kind::Drop &&
           std::find(inter.begin(), inter.end(), queue[end].in[0]) != inter.end()) {
      drops.insert(queue[end].in[0]);
      ++end;
    }
    for (int64_t id : inter)
      if (!drops.count(id)) return none;          // an intermediate is still wanted: leave it to the other fusers
    for (int64_t id : drops)
      if (std::find(inter.begin(), inter.end(), id) == inter.end() && id != xid) return none;
    c.row_in = xid;
    c.dropped = drops;
    c.n_ops = (int)n_ops;
    c.consumed = (int)end;
    return c;
  }

  // ---- Reduce fuser: [read block] → ReduceDim → [write block]
  Candidate scan_reduce() {
    Candidate cur, best;
    cur.kind = B200H_BLOCK_REDUCE;
    size_t q = 0;
    std::vector<int64_t> in_shape;
    for (; q < queue.size(); ++q) {
      const Op &o = queue[q];
      if (o.kind == Kind::Drop) {
        if (cur.read_ops.empty()) return best;
        cur.dropped.insert(o.in[0]);
        continue;
      }
      if (o.kind == Kind::ReduceDim) break;
      if (!is_elemwise(o.kind)) return best;
      const Tensor *out = get(o.out);
      if (cur.read_ops.empty()) in_shape = out->shape;
      else if (out->shape != in_shape) return best;
      cur.read_ops.push_back(&o);
    }
    if (q >= queue.size()) return best;
    const Op &red = queue[q];
    const Tensor *rin = get(red.in[0]);
    if (!cur.read_ops.empty()) {
      // the reduced value must be the last read op's result and the read shape the reduce input shape
      if (cur.read_ops.back()->out != red.in[0] || rin->shape != in_shape) return best;
    }
    cur.core = &red;
    auto evaluate = [&](size_t consumed) {
      Candidate c = cur;
      c.consumed = (int)consumed;
      c.n_ops = (int)(c.read_ops.size() + 1 + c.write_ops.size());
      if (!c.read_ops.empty()) {
        // nothing produced by the read block may need materialising (the kernel has no read outputs)
        std::unordered_set<int64_t> all_dropped = c.dropped;
        for (auto *o : c.read_ops)
          if (!c.dropped.count(o->out)) return;
        c.read_tape = build_tape(c.read_ops, all_dropped, -1, false, 0);
        if (!c.read_tape.ok) return;
      }
      if (!c.write_ops.empty()) {
        c.write_tape = build_tape(c.write_ops, c.dropped, red.out, !c.dropped.count(red.out), B200_MAX_TAPE_OUTPUTS);
        if (!c.write_tape.ok) return;
      }
      best = c;
    };
    evaluate(q + 1);
    const std::vector<int64_t> out_shape = get(red.out)->shape;
    for (size_t w = q + 1; w < queue.size(); ++w) {
      const Op &o = queue[w];
      if (o.kind == Kind::Drop) {
        cur.dropped.insert(o.in[0]);
        evaluate(w + 1);
        continue;
      }
      if (!is_elemwise(o.kind) || get(o.out)->shape != out_shape) break;
      // write-block inputs must be readable at the output shape: reject reads of read-block values
      bool ok = true;
      for (int k = 0; k < 3; ++k)
        for (auto *r : cur.read_ops)
          if (o.in[k] == r->out) ok = false;
      if (!ok) break;
      cur.write_ops.push_back(&o);
      evaluate(w + 1);
    }
    return best;
  }

  // ---- Matmul fuser: Matmul → [epilogue block on the output shape]
  Candidate scan_matmul() {
    Candidate cur, best;
    cur.kind = B200H_BLOCK_MATMUL;
    if (queue.empty() || queue[0].kind != Kind::Matmul) return best;
    const Op &mm = queue[0];
    cur.core = &mm;
    cur.consumed = 1;
    cur.n_ops = 1;
    best = cur;
    const std::vector<int64_t> out_shape = get(mm.out)->shape;
    if (out_shape.back() % 4 != 0) return best;  // the fused epilogue needs N % 4 == 0
    for (size_t w = 1; w < queue.size(); ++w) {
      const Op &o = queue[w];
      if (o.kind == Kind::Drop) {
        cur.dropped.insert(o.in[0]);
      } else {
        if (!is_elemwise(o.kind) || get(o.out)->shape != out_shape) break;
        cur.write_ops.push_back(&o);
      }
      if (cur.write_ops.empty()) { best.consumed = (int)w + 1; best.dropped = cur.dropped; continue; }
      if (!cur.dropped.count(mm.out)) continue;  // the raw product would have to be written too
      TapeBuild tb = build_tape(cur.write_ops, cur.dropped, mm.out, false, 1);
      if (!tb.ok || tb.outputs.size() != 1) continue;
      best = cur;
      best.consumed = (int)w + 1;
      best.n_ops = 1 + (int)cur.write_ops.size();
      best.write_tape = tb;
    }
    if (best.write_ops.empty()) { best.dropped.clear(); best.consumed = 1; }
    return best;
  }

  // ---- descriptors
  int32_t desc_of(int64_t id, b200_tensor &d) {
    Tensor *t = get(id);
    B200_REQUIRE(t, B200_ERR_INVALID, "unknown tensor id %lld", (long long)id);
    B200_REQUIRE(t->buf || plan_only, B200_ERR_INVALID, "tensor %lld has no storage yet", (long long)id);
    memset(&d, 0, sizeof(d));
    d.ptr = t->buf ? (char *)t->buf->ptr + t->offset * b200::dtype_size(t->dtype) : nullptr;
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
      b->ptr = nullptr;
    } else {
      int32_t st = b200_alloc(&b->ptr, (uint64_t)std::max<int64_t>(numel(t->shape), 1) * b200::dtype_size(t->dtype), nullptr);
      if (st != B200_OK) return st;
    }
    t->buf = b;
    t->strides = contiguous(t->shape);
    t->offset = 0;
    return B200_OK;
  }

  static b200_tape as_tape(const TapeBuild &tb) {
    b200_tape t;
    t.ops = tb.ops.data();
    t.n_ops = (int32_t)tb.ops.size();
    t.scalars = tb.scalars.empty() ? nullptr : tb.scalars.data();
    t.n_scalars = (int32_t)tb.scalars.size();
    return t;
  }

  int32_t gather_descs(const std::vector<int64_t> &ids, size_t skip, std::vector<b200_tensor> &out) {
    out.clear();
    for (size_t i = skip; i < ids.size(); ++i) {
      b200_tensor d;
      int32_t st = desc_of(ids[i], d);
      if (st != B200_OK) return st;
      out.push_back(d);
    }
    return B200_OK;
  }

  int32_t execute(const Candidate &c) {
    b200h_block_info info = {c.kind, c.n_ops, 0, 0, 0, 0};
    int32_t st = B200_OK;
    std::vector<b200_tensor> ins, outs, wins;
    const uint64_t before = b200_launch_count();
    if (c.kind == B200H_BLOCK_ELEMWISE) {
      for (auto id : c.read_tape.outputs) if ((st = materialise(id)) != B200_OK) return st;
      if ((st = gather_descs(c.read_tape.inputs, 0, ins)) != B200_OK) return st;
      if ((st = gather_descs(c.read_tape.outputs, 0, outs)) != B200_OK) return st;
      info.n_inputs = (int)ins.size();
      info.n_outputs = (int)outs.size();
      info.n_tape_ops = (int)c.read_tape.ops.size();
      if (!plan_only && !outs.empty()) {
        const Tensor *ref = get(c.read_ops[0]->out);
        b200_tape tape = as_tape(c.read_tape);
        st = b200_launch_elemwise(&tape, ins.data(), (int)ins.size(), outs.data(), (int)outs.size(),
                                  (int)ref->shape.size(), ref->shape.data(), nullptr);
      }
    } else if (c.kind == B200H_BLOCK_REDUCE) {
      const Op &red = *c.core;
      std::vector<int64_t> out_ids = c.write_ops.empty() ? std::vector<int64_t>{red.out} : c.write_tape.outputs;
      for (auto id : out_ids) if ((st = materialise(id)) != B200_OK) return st;
      std::vector<int64_t> in_ids = c.read_ops.empty() ? std::vector<int64_t>{red.in[0]} : c.read_tape.inputs;
      if ((st = gather_descs(in_ids, 0, ins)) != B200_OK) return st;
      if (!c.write_ops.empty() && (st = gather_descs(c.write_tape.inputs, 1, wins)) != B200_OK) return st;
      if ((st = gather_descs(out_ids, 0, outs)) != B200_OK) return st;
      info.n_inputs = (int)(ins.size() + wins.size());
      info.n_outputs = (int)outs.size();
      info.n_tape_ops = (int)(c.read_tape.ops.size() + c.write_tape.ops.size());
      if (!plan_only) {
        const Tensor *rin = get(red.in[0]);
        b200_tape rt = as_tape(c.read_tape), wt = as_tape(c.write_tape);
        st = b200_launch_reduce(red.opcode, red.dim, (int)rin->shape.size(), rin->shape.data(),
                                c.read_ops.empty() ? nullptr : &rt, ins.data(), (int)ins.size(),
                                c.write_ops.empty() ? nullptr : &wt, wins.empty() ? nullptr : wins.data(),
                                (int)wins.size(), outs.data(), (int)outs.size(), nullptr);
      }
    } else if (c.kind == B200H_BLOCK_ROWNORM) {
      if ((st = materialise(c.row_out)) != B200_OK) return st;
      b200_tensor xin, yout;
      if ((st = desc_of(c.row_in, xin)) != B200_OK || (st = desc_of(c.row_out, yout)) != B200_OK) return st;
      info.n_inputs = 1 + (c.row_gamma >= 0) + (c.row_beta >= 0);
      info.n_outputs = 1;
      if (c.row_kind == 2) {
        b200_tensor g, bt;
        if ((st = desc_of(c.row_gamma, g)) != B200_OK) return st;
        if (c.row_beta >= 0 && (st = desc_of(c.row_beta, bt)) != B200_OK) return st;
        if (!plan_only) st = b200_launch_layer_norm(&xin, &g, c.row_beta >= 0 ? &bt : nullptr, c.row_eps, &yout, nullptr);
      } else if (!plan_only) {
        st = b200_launch_softmax(&xin, &yout, c.row_kind == 1 ? 1 : 0, nullptr);
      }
    } else if (c.kind == B200H_BLOCK_MATMUL) {
      const Op &mm = *c.core;
      const int64_t out_id = c.write_ops.empty() ? mm.out : c.write_tape.outputs[0];
      if ((st = materialise(out_id)) != B200_OK) return st;
      b200_tensor a, b, cc;
      if ((st = desc_of(mm.in[0], a)) != B200_OK || (st = desc_of(mm.in[1], b)) != B200_OK ||
          (st = desc_of(out_id, cc)) != B200_OK)
        return st;
      if (!c.write_ops.empty() && (st = gather_descs(c.write_tape.inputs, 1, wins)) != B200_OK) return st;
      info.n_inputs = 2 + (int)wins.size();
      info.n_outputs = 1;
      info.n_tape_ops = (int)c.write_tape.ops.size();
      if (!plan_only) {
        uint64_t wsb = 0;
        if ((st = b200_matmul_workspace_bytes(&a, &b, mm.precision, &wsb)) != B200_OK) return st;
        void *ws = nullptr;
        if (wsb && (st = b200_alloc(&ws, wsb, nullptr)) != B200_OK) return st;
        b200_tape wt = as_tape(c.write_tape);
        st = b200_launch_matmul(&a, &b, &cc, mm.precision, c.write_ops.empty() ? nullptr : &wt,
                                wins.empty() ? nullptr : wins.data(), (int)wins.size(), ws, wsb, nullptr);
        if (ws) b200_free(ws, nullptr);
      }
    }
    if (st != B200_OK) return st;
    info.launches = (int32_t)(b200_launch_count() - before);
    log.push_back(info);
    for (auto id : c.dropped) tensors.erase(id);  // R::free_handle for consumed ReadWrite handles
    return B200_OK;
  }

  int32_t drain() {
    while (!queue.empty()) {
      if (queue[0].kind == Kind::Drop) {
        tensors.erase(queue[0].in[0]);
        queue.erase(queue.begin());
        continue;
      }
      Candidate best = scan_elemwise();
      Candidate r = scan_reduce();
      Candidate m = scan_matmul();
      if (r.n_ops > best.n_ops || (r.n_ops == best.n_ops && r.n_ops > 0)) best = r;
      if (m.n_ops >= best.n_ops && m.n_ops > 0) best = m;
      Candidate rn = scan_rownorm();
      if (rn.n_ops > best.n_ops) best = rn;
      B200_REQUIRE(best.n_ops > 0, B200_ERR_UNSUPPORTED, "no fuser accepts operation kind %d", (int)queue[0].kind);
      // ops never produce tensors that need materialising? pending outputs of dropped ids vanish
      int32_t st = execute(best);
      if (st != B200_OK) return st;
      queue.erase(queue.begin(), queue.begin() + best.consumed);
    }
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

int32_t b200h_sync(b200h_stream s) {
  int32_t st = S(s)->drain();
  if (st != B200_OK) return st;
  return S(s)->plan_only ? B200_OK : b200_stream_sync(nullptr);
}

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

b200h_id b200h_swap_dims(b200h_stream s, b200h_id x, int32_t d0, int32_t d1) {
  // metadata-only view (crates/burn-cubecl/src/ops/base.rs:137-139); a queued producer is drained first
  NEED(s, x, probe);
  (void)probe;
  if (!S(s)->get(x)->buf && S(s)->drain() != B200_OK) return -1;
  Tensor *t = S(s)->get(x);
  const int r = (int)t->shape.size();
  if (d0 < 0) d0 += r;
  if (d1 < 0) d1 += r;
  if (d0 < 0 || d1 < 0 || d0 >= r || d1 >= r) { b200::fail(B200_ERR_SHAPE, "swap_dims out of range"); return -1; }
  Tensor v = *t;
  std::swap(v.shape[d0], v.shape[d1]);
  std::swap(v.strides[d0], v.strides[d1]);
  const int64_t id = S(s)->next_id++;
  S(s)->tensors[id] = std::move(v);
  return id;
}

int32_t b200h_drop(b200h_stream s, b200h_id id) {
  B200_REQUIRE(S(s)->get(id), B200_ERR_INVALID, "unknown tensor id %lld", (long long)id);
  Op o;
  o.kind = Kind::Drop;
  o.in[0] = id;
  S(s)->queue.push_back(o);
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
