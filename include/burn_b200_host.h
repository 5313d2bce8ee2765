/*
 * burn_b200_host.h — C face of the host-side fusion layer (burn_b200/host/fusion.cpp).
 *
 * The layer is the C++ stand-in for the Rust code a `crates/burn-b200` would contain above the
 * kernel ABI (include/burn_b200.h): it receives a lazy stream of operations shaped like
 * burn-ir's `OperationIr` (crates/burn-ir/src/operation.rs:113-142), runs the three
 * `OperationFuser` state machines the reference registers — ElementWise, Matmul, Reduce
 * (crates/burn-cubecl/src/fusion/registry.rs:128-146; acceptance rules
 * crates/burn-cubecl-fusion/src/engine/fuser.rs:76-190,292-710,
 * crates/burn-cubecl-fusion/src/optim/{reduce,matmul}/fuser.rs)
 * (plus the ReduceBroadcasted shape: the max_dim/sub/exp/sum_dim/div[/log/sub] chains of
 * crates/burn-backend/src/backend/ops/activation.rs:250-276 → b200_launch_softmax,
 * crates/burn-cubecl-fusion/src/optim/reduce_broadcasted/) — and executes each fused block
 * as ONE kernel through b200_launch_elemwise / b200_launch_reduce / b200_launch_matmul
 * (`Optimization::execute`, crates/burn-fusion/src/backend.rs:226-234).
 *
 * Tensors are named by integer ids like `TensorIr::id`; `b200h_drop` is `OperationIr::Drop`
 * (it is how intermediates stay in registers, fuser.rs:87-95).  Everything is lazy until
 * `b200h_read` / `b200h_sync` drain the stream (crates/burn-fusion/src/client.rs:153-201).
 * With plan_only = 1 no device is touched: blocks are planned and logged, which is how the
 * CPU tests check fusion decisions (the reference does this with fake backends,
 * crates/burn-fusion/src/stream/execution/tests.rs).
 */
#ifndef BURN_B200_HOST_H
#define BURN_B200_HOST_H

#include <stdint.h>

#include "burn_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef void *b200h_stream;
typedef int64_t b200h_id; /* < 0 = error (see b200_last_error) */

typedef enum {
  B200H_BLOCK_ELEMWISE = 0,
  B200H_BLOCK_REDUCE = 1,
  B200H_BLOCK_MATMUL = 2,
  B200H_BLOCK_EAGER = 3, /* op no fuser accepts, executed on its own */
  B200H_BLOCK_ROWNORM = 4, /* softmax / log_softmax / layer_norm chain → one row-resident kernel (ReduceBroadcasted) */
  B200H_BLOCK_VIEW = 5     /* a lone reshape / swap_dims / expand / slice of a still-pending tensor: metadata only, no launch
                              (crates/burn-backend-tests/tests/fusion/fusion_shape.rs:233 lone_view_is_not_fused_into_kernel) */
} b200h_block_kind;

/* One executed (or planned) optimization — the FusionInspector view
 * (crates/burn-fusion/src/inspect.rs:1-45). */
typedef struct {
  int32_t kind;      /* b200h_block_kind */
  int32_t n_ops;     /* IR operations absorbed (Drop not counted) */
  int32_t n_inputs;  /* global tensors read */
  int32_t n_outputs; /* global tensors written */
  int32_t n_tape_ops;/* public tape ops handed to the kernel (read+write tapes summed) */
  int32_t launches;  /* kernel launches issued for the block */
  int32_t aliased;   /* outputs written in place over a consumed (ReadWrite) input: HandleOutput::Alias,
                        crates/burn-cubecl-fusion/src/engine/launch/output.rs:47-55 */
  int32_t from_cache;/* 1 = the block came from the plan cache (no fuser ran for it) */
  uint64_t score;    /* FuserProperties::score of the winning fuser (scoring.rs:56-76) */
} b200h_block_info;

int32_t b200h_stream_create(b200h_stream *out, int32_t plan_only);
/* Drain the pending queue (plan + launch every block) without waiting for the device: what a graph capture or an
 * asynchronous caller uses; b200h_sync = flush + stream synchronize. */
int32_t b200h_flush(b200h_stream s);
int32_t b200h_stream_destroy(b200h_stream s);

/* float_from_data / bool_from_data: contiguous host data → new tensor id. */
b200h_id b200h_from_host(b200h_stream s, const void *data, int32_t dtype, int32_t rank, const int64_t *shape);
/* float_into_data: drains the stream, copies the (contiguous) tensor to `dst`. */
int32_t b200h_read(b200h_stream s, b200h_id id, void *dst, uint64_t dst_bytes);
int32_t b200h_shape(b200h_stream s, b200h_id id, int32_t *dtype, int32_t *rank, int64_t *shape);
int32_t b200h_sync(b200h_stream s); /* drain without reading */

/* Lazy operations; each returns the id of its output tensor. `opcode` is a b200_opcode. */
b200h_id b200h_binary(b200h_stream s, int32_t opcode, b200h_id lhs, b200h_id rhs);       /* Add … Lower … */
b200h_id b200h_scalar(b200h_stream s, int32_t opcode, b200h_id lhs, double scalar);      /* AddScalar … LowerElem … */
b200h_id b200h_unary(b200h_stream s, int32_t opcode, b200h_id x);                        /* Exp, Erf, Sqrt … */
b200h_id b200h_mask_fill(b200h_stream s, b200h_id x, b200h_id mask, double value);
b200h_id b200h_mask_where(b200h_stream s, b200h_id x, b200h_id mask, b200h_id source);
b200h_id b200h_reduce_dim(b200h_stream s, int32_t kind, b200h_id x, int32_t dim);        /* SumDim, MeanDim, ArgMax … (keepdim) */
b200h_id b200h_matmul(b200h_stream s, b200h_id lhs, b200h_id rhs, int32_t precision);
b200h_id b200h_swap_dims(b200h_stream s, b200h_id x, int32_t d0, int32_t d1);            /* metadata-only view */
/* BaseOperationIr::{Reshape, Expand, Slice} (crates/burn-ir/src/operation.rs): metadata only.  On a tensor that already
 * has storage the view is resolved at once; on a pending one it is queued as a lone view block (it closes the fusers,
 * as in the reference where a view is fused only as an INPUT of a later block).  A reshape of a non-contiguous tensor
 * first copies (into_contiguous).  Slice bounds are already canonical (0 <= start <= end <= dim), unit steps. */
b200h_id b200h_reshape(b200h_stream s, b200h_id x, int32_t rank, const int64_t *shape);
b200h_id b200h_expand(b200h_stream s, b200h_id x, int32_t rank, const int64_t *shape);
b200h_id b200h_slice(b200h_stream s, b200h_id x, const int64_t *starts, const int64_t *ends);
/* NumericOperationIr::{Gather, Select} (fuser.rs:381-420 fuses them as indexed reads; here they run as their own
 * block through b200_launch_gather / b200_launch_select — stated gap). */
b200h_id b200h_gather(b200h_stream s, int32_t dim, b200h_id x, b200h_id indices);
b200h_id b200h_select(b200h_stream s, int32_t dim, b200h_id x, b200h_id indices);
int32_t b200h_drop(b200h_stream s, b200h_id id);                                          /* OperationIr::Drop */
/* Wrap device memory the caller owns (non-owning handle) / look a tensor's storage up after a drain: how kernels
 * outside the stream (attention, optimizer) exchange tensors with it. */
b200h_id b200h_from_device(b200h_stream s, void *ptr, int32_t dtype, int32_t rank, const int64_t *shape, const int64_t *strides);
int32_t b200h_device_tensor(b200h_stream s, b200h_id id, b200_tensor *out);               /* drains first */

/* ---- plan cache (crates/burn-fusion/src/stream/store: ExecutionPlanStore).  A drained queue is first rewritten in
 * RELATIVE form — tensor ids renumbered by first appearance, every distinct dimension value replaced by a relative
 * shape id (1 is always id 0), scalars and view arguments lifted out into the Context
 * (crates/burn-fusion/src/stream/context.rs:11-26,56-90) — and the fusers' decisions are cached under that form.
 * The same op sequence at other concrete sizes / scalar values re-executes the cached optimizations with a new
 * Context; no fuser runs. */
typedef struct {
  uint64_t hits, misses;    /* drains served from / added to the cache */
  uint64_t plans;           /* cached relative traces */
  uint64_t inplace_aliases; /* process-wide count of outputs written in place (inspect::inplace_alias_count) */
} b200h_cache_stats;
int32_t b200h_cache_stats_get(b200h_stream s, b200h_cache_stats *out);
int32_t b200h_cache_clear(b200h_stream s);

/* ---- the OperationFuser / Optimization contract itself (crates/burn-fusion/src/backend.rs:157-234), over the
 * stream's pending queue as the operation source.  The drain loop above is written on exactly these objects; they are
 * exported so the boundary can be driven (and tested) the way burn-fusion's Processor / Explorer drives it:
 * feed OperationIr one at a time to every fuser until all are Closed, compare properties(), clone_dyn() a fuser to
 * explore an alternative, finish() the winner, execute() it — or to_state() it and from_state() it elsewhere. */
typedef void *b200h_fuser;
typedef void *b200h_optimization;
typedef enum { B200H_FUSER_OPEN = 0, B200H_FUSER_CLOSED = 1 } b200h_fuser_status;
int32_t b200h_fuser_create(b200h_stream s, int32_t block_kind, b200h_fuser *out); /* ELEMWISE / REDUCE / MATMUL / ROWNORM */
int32_t b200h_fuser_destroy(b200h_fuser f);
int32_t b200h_fuser_fuse_next(b200h_fuser f);              /* fuse(&operation): registers the next queued operation */
int32_t b200h_fuser_status_get(b200h_fuser f);             /* b200h_fuser_status */
int32_t b200h_fuser_properties(b200h_fuser f, uint64_t *score, int32_t *ready);
int32_t b200h_fuser_len(b200h_fuser f);                    /* operations fused so far (Drop not counted) */
int32_t b200h_fuser_reset(b200h_fuser f);
int32_t b200h_fuser_clone(b200h_fuser f, b200h_fuser *out);/* clone_dyn */
int32_t b200h_fuser_finish(b200h_fuser f, b200h_optimization *out);
int32_t b200h_optimization_destroy(b200h_optimization o);
int32_t b200h_optimization_len(b200h_optimization o);      /* NumOperations::len */
const char *b200h_optimization_name(b200h_optimization o); /* NumOperations::name */
/* execute(&mut self, context, ..): binds the optimization's relative tensors / scalars to the head of `s`'s pending
 * queue (which must relativise to the operations the optimization was built from — B200_ERR_INVALID otherwise),
 * launches ONE kernel, registers the outputs, frees consumed handles, and pops those operations. */
int32_t b200h_optimization_execute(b200h_optimization o, b200h_stream s);
/* to_state / from_state: a self-contained byte string (relative ids only — no pointers, no shapes). */
int32_t b200h_optimization_to_state(b200h_optimization o, void *buf, uint64_t cap, uint64_t *len);
int32_t b200h_optimization_from_state(const void *buf, uint64_t len, b200h_optimization *out);

/* Inspector: blocks executed since creation / last clear. */
int32_t b200h_block_count(b200h_stream s);
int32_t b200h_block_get(b200h_stream s, int32_t index, b200h_block_info *out);
int32_t b200h_block_clear(b200h_stream s);

#ifdef __cplusplus
}
#endif
#endif
