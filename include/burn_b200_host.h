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
  B200H_BLOCK_ROWNORM = 4 /* softmax / log_softmax / layer_norm chain → one row-resident kernel (ReduceBroadcasted) */
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
} b200h_block_info;

int32_t b200h_stream_create(b200h_stream *out, int32_t plan_only);
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
int32_t b200h_drop(b200h_stream s, b200h_id id);                                          /* OperationIr::Drop */

/* Inspector: blocks executed since creation / last clear. */
int32_t b200h_block_count(b200h_stream s);
int32_t b200h_block_get(b200h_stream s, int32_t index, b200h_block_info *out);
int32_t b200h_block_clear(b200h_stream s);

#ifdef __cplusplus
}
#endif
#endif
