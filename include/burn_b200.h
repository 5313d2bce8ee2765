/*
 * burn_b200.h — C ABI of libburn_b200.so, the sm_100a CUDA library behind the
 * `burn-b200` backend (Burn `Backend` trait + burn-fusion `FusionRuntime`).
 *
 * This is the drop-in boundary: every entry point below is what the Rust crate
 * `crates/burn-b200` binds with `extern "C"` (see INTEGRATION.md for the Rust
 * declarations).  Signatures carry only plain pointers, sizes and PODs.  Each
 * entry point cites the reference interface it serves (paths relative to the
 * tracel-ai/burn tree, version 0.22.0-pre.2).
 *
 * Conventions
 *  - every function returns 0 on success, a negative b200_status otherwise;
 *    b200_last_error() returns a thread-local message.  The Rust shim turns a
 *    non-zero status into panic!/ExecutionError exactly where the reference
 *    panics (crates/burn-backend/src/backend/ops/tensor.rs — shape/dtype errors
 *    panic; only sync / into_data return Result).
 *  - tensors are described by b200_tensor: device pointer, dtype, rank, shape
 *    and strides IN ELEMENTS (stride 0 = broadcast).  This mirrors
 *    CubeTensor{handle, meta(shape+strides), dtype}
 *    (crates/burn-cubecl/src/tensor/base.rs:20-33).
 *  - all launches are asynchronous on the given b200_stream (NULL = the
 *    library's per-device default stream).
 *  - there is no CPU fallback anywhere in this library.
 */
#ifndef BURN_B200_H
#define BURN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_ABI_VERSION 1
#define B200_MAX_RANK 8
#define B200_MAX_TAPE_OPS 64     /* crates/burn-cubecl-fusion/src/engine/fuser.rs:857 */
#define B200_MAX_TAPE_INPUTS 12
#define B200_MAX_TAPE_OUTPUTS 8
#define B200_MAX_TAPE_TEMPS 8
#define B200_MAX_TAPE_SCALARS 32

typedef enum {
  B200_OK = 0,
  B200_ERR_CUDA = -1,        /* a CUDA runtime/driver call failed */
  B200_ERR_INVALID = -2,     /* bad argument (rank, dtype, null pointer …) */
  B200_ERR_SHAPE = -3,       /* shape mismatch — the reference panics here */
  B200_ERR_UNSUPPORTED = -4, /* valid request this library does not implement */
  B200_ERR_NCCL = -5,
  B200_ERR_NO_DEVICE = -6
} b200_status;

/* DType subset of crates/burn-std/src/tensor/dtype.rs:10-26.  Bool is stored as
 * one byte per element (BoolStore::U8 — the CUDA default,
 * crates/burn-backend/src/lib.rs:98-107). */
typedef enum {
  B200_F32 = 0,
  B200_F16 = 1,
  B200_BF16 = 2,
  B200_I32 = 3,
  B200_I64 = 4,
  B200_BOOL = 5, /* u8, 0 or 1 */
  B200_U8 = 6,
  B200_DTYPE_COUNT = 7
} b200_dtype;

typedef struct {
  void *ptr;
  int32_t dtype; /* b200_dtype */
  int32_t rank;
  int64_t shape[B200_MAX_RANK];
  int64_t strides[B200_MAX_RANK]; /* in elements; 0 = broadcast */
} b200_tensor;

typedef void *b200_stream; /* cudaStream_t */

/* ------------------------------------------------------------------ runtime */
/* Backend::{name, device_count, sync, seed, memory_cleanup}
 * (crates/burn-backend/src/backend/base.rs:112-250). */
int32_t b200_abi_version(void);
int32_t b200_device_count(int32_t *count);
int32_t b200_init(int32_t device);              /* idempotent; selects device */
int32_t b200_set_device(int32_t device);
int32_t b200_device_info(int32_t device, int32_t *sm_count, int32_t *cc_major,
                         int32_t *cc_minor, uint64_t *total_mem);
int32_t b200_stream_create(b200_stream *out, int32_t high_priority);
int32_t b200_stream_destroy(b200_stream s);
/* Both also report (B200_ERR_SHAPE) an out-of-range index seen by an earlier gather / select /
 * scatter_add / select_add / cross-entropy launch: the reference panics on those
 * (crates/burn-ndarray/src/ops/base.rs:106-183), a kernel cannot, so the access is skipped and the error
 * surfaces at the next synchronising call. */
int32_t b200_stream_sync(b200_stream s);        /* Backend::sync */
int32_t b200_device_sync(void);
const char *b200_last_error(void);
/* CUDA events on the launching stream — what bench.py times kernels with. */
typedef void *b200_event;
int32_t b200_event_create(b200_event *out);
int32_t b200_event_destroy(b200_event e);
int32_t b200_event_record(b200_event e, b200_stream s);
int32_t b200_event_query(b200_event e, int32_t *done);   /* async into_data: poll instead of blocking */
int32_t b200_event_elapsed_ms(b200_event start, b200_event stop, float *ms); /* syncs on stop */

/* CUDA-graph capture of a launch sequence (the replacement
 * for re-walking burn-fusion's ExecutionPlan store every step,
 * crates/burn-fusion/src/stream/store/base.rs — a training step whose shapes do
 * not change is captured once and replayed with one driver call).  Everything
 * issued on `s` between begin and end — kernels, b200_alloc/b200_free (become
 * graph memory nodes), memsets, d2d copies, and collectives forked onto the comm
 * stream and joined back by b200_collective_sync — is recorded, not executed.
 * Buffers allocated before the capture keep their addresses across replays;
 * buffers allocated inside it must also be freed inside it.  A graph that
 * captured collectives must be destroyed before b200_comm_destroy (NCCL waits
 * for every graph that references the communicator). */
typedef void *b200_graph;
int32_t b200_graph_begin(b200_stream s);
int32_t b200_graph_end(b200_stream s, b200_graph *out);
int32_t b200_graph_launch(b200_graph g, b200_stream s);
int32_t b200_graph_destroy(b200_graph g);
int32_t b200_graph_node_count(b200_graph g, uint64_t *kernel_nodes, uint64_t *total_nodes);

/* Stream-ordered caching allocator (cudaMallocAsync pool, never trimmed until
 * b200_memory_cleanup).  Replaces cubecl's memory pools behind
 * CubeTensor::handle; b200_retain/b200_free give the refcount semantics of
 * Handle::can_mut (crates/burn-ir/src/handle.rs:92-111). */
int32_t b200_alloc(void **out, uint64_t bytes, b200_stream s);   /* refcount 1 */
int32_t b200_retain(void *ptr);                                   /* +1 owner (tensor clone = refcount bump) */
int32_t b200_refcount(const void *ptr, uint32_t *count);          /* count == 1  <=>  Handle::can_mut */
int32_t b200_free(void *ptr, b200_stream s);                      /* -1 owner; memory released at 0 */
int32_t b200_memory_cleanup(void);
int32_t b200_memset(void *ptr, int32_t byte, uint64_t bytes, b200_stream s);

/* float_from_data / float_into_data / float_to_device
 * (crates/burn-backend/src/backend/ops/tensor.rs:25,109,123).  Host buffers
 * may be pageable; pinned buffers (b200_host_alloc) make the copy truly async. */
int32_t b200_host_alloc(void **out, uint64_t bytes);
int32_t b200_host_free(void *ptr);
int32_t b200_memcpy_h2d(void *dst, const void *src, uint64_t bytes, b200_stream s);
int32_t b200_memcpy_d2h(void *dst, const void *src, uint64_t bytes, b200_stream s);
int32_t b200_memcpy_d2d(void *dst, const void *src, uint64_t bytes, b200_stream s);

/* ---------------------------------------------------- path (a): op tapes */
/*
 * A tape is the register-resident program one fused kernel runs per element.
 * It is what `TraceOperationFuser` builds as Vec<FuseOp>
 * (crates/burn-cubecl-fusion/src/engine/fuser.rs:33-839,
 *  crates/burn-cubecl-fusion/src/engine/codegen/ir.rs:135-195), flattened to an
 * accumulator machine: every op writes the accumulator (ACC); it may also save
 * the result into a temp slot and/or store it to an output tensor.
 *
 * Operand byte: bits 7..6 kind, bits 5..0 index.
 */
#define B200_ARG_ACC 0x00u             /* result of the previous tape op */
#define B200_ARG_INPUT(i) (0x40u | (uint8_t)(i))
#define B200_ARG_TEMP(i) (0x80u | (uint8_t)(i))
#define B200_ARG_SCALAR(i) (0xC0u | (uint8_t)(i))
#define B200_DST_NONE 0xFFu

typedef enum {
  /* float (f32 math; f16/bf16 are storage types converted at load/store) */
  B200_OP_MOV = 0, /* FuseOp::Assign */
  B200_OP_ADD_F, B200_OP_SUB_F, B200_OP_MUL_F, B200_OP_DIV_F, B200_OP_REM_F,
  B200_OP_POW_F, B200_OP_MIN_F, B200_OP_MAX_F, B200_OP_ATAN2_F,
  B200_OP_NEG_F, B200_OP_ABS_F, B200_OP_EXP_F, B200_OP_LOG_F, B200_OP_LOG1P_F,
  B200_OP_SQRT_F, B200_OP_RECIP_F, B200_OP_SIN_F, B200_OP_COS_F, B200_OP_TAN_F,
  B200_OP_TANH_F, B200_OP_ERF_F, B200_OP_FLOOR_F, B200_OP_CEIL_F,
  B200_OP_ROUND_F, B200_OP_TRUNC_F, B200_OP_SIGN_F, B200_OP_SINH_F,
  B200_OP_COSH_F, B200_OP_ASIN_F, B200_OP_ACOS_F, B200_OP_ATAN_F,
  B200_OP_ASINH_F, B200_OP_ACOSH_F, B200_OP_ATANH_F, B200_OP_SIGMOID_F,
  B200_OP_CLAMP_F, /* a, lo=b, hi=c */
  /* float comparisons -> bool */
  B200_OP_EQ_F, B200_OP_NE_F, B200_OP_LT_F, B200_OP_LE_F, B200_OP_GT_F,
  B200_OP_GE_F, B200_OP_ISNAN_F, B200_OP_ISINF_F,
  /* int (i32 math; i64 is a storage type) */
  B200_OP_ADD_I, B200_OP_SUB_I, B200_OP_MUL_I, B200_OP_DIV_I, B200_OP_REM_I,
  B200_OP_MIN_I, B200_OP_MAX_I, B200_OP_NEG_I, B200_OP_ABS_I, B200_OP_SIGN_I,
  B200_OP_AND_I, B200_OP_OR_I, B200_OP_XOR_I, B200_OP_NOT_I, B200_OP_SHL_I,
  B200_OP_SHR_I, B200_OP_CLAMP_I,
  B200_OP_EQ_I, B200_OP_NE_I, B200_OP_LT_I, B200_OP_LE_I, B200_OP_GT_I,
  B200_OP_GE_I,
  /* bool */
  B200_OP_AND_B, B200_OP_OR_B, B200_OP_XOR_B, B200_OP_NOT_B,
  /* select: out = cond(c) ? b : a  — FuseOp::ConditionalAssign; mask_fill is
   * a=input, b=scalar, c=mask; mask_where is a=input, b=source, c=mask
   * (crates/burn-ndarray/src/ops/base.rs:78-104). */
  B200_OP_SELECT,
  /* casts (FuseOp::Assign with a dtype change; Rust `as` semantics) */
  B200_OP_F2I, B200_OP_I2F, B200_OP_B2F, B200_OP_B2I, B200_OP_F2B, B200_OP_I2B,
  /* tensor-tensor float remainder: a - b*floor(a/b) evaluated in f64 (crates/burn-ndarray/src/ops/base.rs:909-922).
   * The REM_F opcode above is remainder_scalar's ((x % y) + y) % y (base.rs:924-930); the two differ in rounding, in the sign of
   * zero and for infinite divisors.  Appended so that the existing opcode values stay what they were. */
  B200_OP_REMT_F,
  B200_OP_COUNT
} b200_opcode;

typedef struct {
  uint8_t op;      /* b200_opcode */
  uint8_t a, b, c; /* operand bytes */
  uint8_t dst_temp;/* temp slot to save the result into, or B200_DST_NONE */
  uint8_t dst_out; /* output tensor index to store the result to, or NONE */
  uint8_t pad[2];
} b200_tape_op;

typedef struct {
  const b200_tape_op *ops;
  int32_t n_ops;
  const uint32_t *scalars; /* raw 32-bit patterns (f32 or i32) */
  int32_t n_scalars;
} b200_tape;

/*
 * Fused elementwise launch — ElemwiseOptimization::execute →
 * elemwise_fuse::launch_unchecked
 * (crates/burn-cubecl-fusion/src/optim/elemwise/optimization.rs:81-171).
 * `ref_shape` is the block's output shape; every input/output descriptor must
 * have rank == `rank` and be broadcast-compatible with it (size-1 dims are
 * given stride 0 by the caller or by the library).  Outputs may alias inputs
 * (in-place reuse, crates/burn-cubecl-fusion/src/engine/launch/output.rs:47-55)
 * when they have the same layout.
 */
int32_t b200_launch_elemwise(const b200_tape *tape, const b200_tensor *inputs,
                             int32_t n_inputs, const b200_tensor *outputs,
                             int32_t n_outputs, int32_t rank,
                             const int64_t *ref_shape, b200_stream s);

/* ------------------------------------------------ path (b): reductions */
/* ReduceDimOpIr{input,out,axis} (crates/burn-ir/src/operation.rs:1008-1020);
 * kinds accepted by the Reduce fuser
 * (crates/burn-cubecl-fusion/src/optim/reduce/fuser.rs:221-300). */
typedef enum {
  B200_RED_SUM = 0,
  B200_RED_MEAN = 1,
  B200_RED_PROD = 2,
  B200_RED_MAX = 3,
  B200_RED_MIN = 4,
  B200_RED_ARGMAX = 5, /* first max, first NaN wins: burn-ndarray base.rs:1715-1757 */
  B200_RED_ARGMIN = 6,
  B200_RED_MAXABS = 7,
  B200_RED_ANY = 8,
  B200_RED_ALL = 9
} b200_reduce_kind;

/*
 * Fused reduce: out = write_tape( reduce_axis( read_tape(inputs…) ) ), keepdim.
 * reduce_kernel_fused (crates/burn-cubecl-fusion/src/optim/reduce/optimization.rs:503).
 *  - `read`  (optional, may be NULL): fuse-on-read tape over `inputs`, evaluated
 *    at the reduce INPUT shape `in_shape`; its last ACC value is what is reduced.
 *    With read == NULL the single input tensor inputs[0] is reduced directly.
 *  - `write` (optional): fuse-on-write tape evaluated at the OUTPUT shape
 *    (in_shape with in_shape[axis] = 1); its INPUT(0) is the reduced value and
 *    INPUT(1…) are `write_inputs`; results go to `outputs`.  With write == NULL
 *    the reduced value is stored to outputs[0].
 *  Float reductions accumulate in f32; arg reductions write i32 or i64 per the
 *  output dtype.  axis == -1 with rank-1 output shape [1] reduces everything
 *  (float_sum, crates/burn-backend/src/backend/ops/tensor.rs:883).
 */
int32_t b200_launch_reduce(int32_t kind, int32_t axis, int32_t rank,
                           const int64_t *in_shape, const b200_tape *read,
                           const b200_tensor *inputs, int32_t n_inputs,
                           const b200_tape *write,
                           const b200_tensor *write_inputs,
                           int32_t n_write_inputs, const b200_tensor *outputs,
                           int32_t n_outputs, b200_stream s);

/* Full reduction of a tensor to shape [1] (Sum/Mean/Prod/Max/Min/Any/All).
 * float_sum → sum_view (crates/burn-ndarray/src/ops/base.rs:940-943). */
int32_t b200_launch_reduce_full(int32_t kind, const b200_tensor *input,
                                const b200_tensor *output, b200_stream s);

/* ------------------------------------------------ path (c): float_matmul */
typedef enum {
  B200_MM_TF32 = 0,   /* f32 operands fed to tcgen05 kind::tf32, f32 accumulate */
  B200_MM_BF16 = 1,   /* operands rounded to bf16 (RN), kind::f16, f32 accumulate */
  B200_MM_F32X3 = 2   /* 3xTF32 split: near-f32 accuracy on the tensor pipe */
} b200_mm_precision;

/*
 * C[..., M, N] = A[..., M, K] · B[..., K, N] with numpy-style broadcast of the
 * leading dims (crates/burn-ndarray/src/ops/matmul.rs:9-183) and arbitrary
 * (transposed-view) strides on the last two dims of A and B — the NT/TN GEMMs
 * autodiff issues (crates/burn-autodiff/src/ops/tensor.rs:560-616).
 * `epilogue` (optional) is a fuse-on-write tape run on the accumulator before
 * the store — MatmulOptimization (crates/burn-cubecl-fusion/src/optim/matmul/
 * optimization.rs:97-140): INPUT(0) is the accumulator, INPUT(1…) are
 * `epi_inputs` described at the OUTPUT shape (broadcast strides allowed).
 * All of a, b, c must have the same rank >= 2.  c must be row-major contiguous.
 * `workspace` (may be NULL) is scratch for operand conversion; query its size
 * with b200_matmul_workspace_bytes.
 */
int32_t b200_matmul_workspace_bytes(const b200_tensor *a, const b200_tensor *b,
                                    int32_t precision, uint64_t *bytes);
int32_t b200_launch_matmul(const b200_tensor *a, const b200_tensor *b,
                           const b200_tensor *c, int32_t precision,
                           const b200_tape *epilogue,
                           const b200_tensor *epi_inputs, int32_t n_epi_inputs,
                           void *workspace, uint64_t workspace_bytes,
                           b200_stream s);

/* ------------------------------------------------ indexing / data movement */
/* strided → contiguous copy with optional dtype cast (float_cast, into_contiguous;
 * crates/burn-cubecl/src/kernel/contiguous.rs, kernel/cast/base.rs:13). */
int32_t b200_launch_copy(const b200_tensor *src, const b200_tensor *dst, b200_stream s);
/* float_gather / int_gather: out[..i..] = t[.. idx[..i..] ..] along dim
 * (crates/burn-ndarray/src/ops/base.rs:106-138). */
int32_t b200_launch_gather(int32_t dim, const b200_tensor *input,
                           const b200_tensor *indices, const b200_tensor *out,
                           b200_stream s);
/* float_scatter_add: t[.. idx ..] += value, sequential along dim per lane —
 * deterministic and bit-identical to the oracle
 * (crates/burn-cubecl/src/kernel/index/scatter.rs:13-71,
 *  crates/burn-ndarray/src/ops/base.rs:140-183).  `tensor` is updated in place. */
int32_t b200_launch_scatter_add(int32_t dim, const b200_tensor *tensor,
                                const b200_tensor *indices,
                                const b200_tensor *value, b200_stream s);
/* float_select: out = t.index_select(dim, indices[1-D])
 * (crates/burn-backend/src/backend/ops/tensor.rs:518). */
int32_t b200_launch_select(int32_t dim, const b200_tensor *input,
                           const b200_tensor *indices, const b200_tensor *out,
                           b200_stream s);
/* float_select_add: t.index_add(dim, indices, value) in place, deterministic
 * (sequential over indices per lane) — embedding backward
 * (crates/burn-backend/src/backend/ops/modules/base.rs:161-180). */
int32_t b200_launch_select_add(int32_t dim, const b200_tensor *tensor,
                               const b200_tensor *indices,
                               const b200_tensor *value, b200_stream s);
/* float_slice_assign (tensor.rs:592; crates/burn-cubecl/src/kernel/index/slice_assign.rs): tensor[starts..ends] = value,
 * in place on `tensor` (the caller owns it, as the by-value trait signature implies); unit steps. */
int32_t b200_launch_slice_assign(const b200_tensor *tensor, const int64_t *starts,
                                 const int64_t *ends, const b200_tensor *value, b200_stream s);
/* float_cat (tensor.rs:1460): inputs[0..n) concatenated along dim into `out`; empty inputs are skipped. */
int32_t b200_launch_cat(const b200_tensor *inputs, int32_t n, int32_t dim,
                        const b200_tensor *out, b200_stream s);
/* float_repeat_dim (tensor.rs:161; kernel/index/repeat_dim.rs): out = input tiled `times` along dim. */
int32_t b200_launch_repeat_dim(const b200_tensor *input, int32_t dim, int64_t times,
                               const b200_tensor *out, b200_stream s);
/* float_flip (tensor.rs:410; kernel/index/flip.rs): out = input reversed along each of axes[0..n_axes). */
int32_t b200_launch_flip(const b200_tensor *input, const int32_t *axes, int32_t n_axes,
                         const b200_tensor *out, b200_stream s);
/* float_random (crates/burn-backend/src/backend/ops/tensor.rs:39): Philox4x32-10.
 * kind 0 = uniform[lo,hi), 1 = normal(mean=lo,std=hi), 2 = bernoulli(p=lo). */
int32_t b200_launch_random(const b200_tensor *out, int32_t kind, double lo,
                           double hi, uint64_t seed, uint64_t offset,
                           b200_stream s);
/* int_arange (crates/burn-backend/src/backend/ops/int_tensor.rs:1287). */
int32_t b200_launch_arange(const b200_tensor *out, int64_t start, int64_t step,
                           b200_stream s);

/* ------------------------------------------------ row-resident fused chains */
/* softmax / log_softmax along the last axis in one pass (the ReduceBroadcasted
 * analogue, crates/burn-cubecl-fusion/src/optim/reduce_broadcasted/; op chain
 * crates/burn-backend/src/backend/ops/activation.rs:250-287). */
int32_t b200_launch_softmax(const b200_tensor *input, const b200_tensor *out,
                            int32_t log_softmax, b200_stream s);
/* layer_norm over the last axis: (x-mean)/sqrt(var+eps)*gamma+beta
 * (crates/burn-backend/src/backend/ops/modules/base.rs:846-877).  gamma/beta
 * may be NULL. */
int32_t b200_launch_layer_norm(const b200_tensor *input, const b200_tensor *gamma,
                               const b200_tensor *beta, double eps,
                               const b200_tensor *out, b200_stream s);
/* Backward of the two row chains in one pass each (what burn-autodiff's reverse
 * walk over those op chains computes, crates/burn-autodiff/src/ops/tensor.rs —
 * div/exp/sub/sum_dim/mean_dim backward steps).
 * softmax: dx = (dy - sum(dy*y)) * y / div, and 0 where `mask` (bool, broadcast
 * over the leading dims; may be NULL) is set — `div` and `mask` fold in the
 * score scaling and mask_fill backward of MultiHeadAttention
 * (crates/burn-nn/src/modules/attention/mha.rs:253-311). */
int32_t b200_launch_softmax_backward(const b200_tensor *y, const b200_tensor *dy,
                                     const b200_tensor *mask, double div,
                                     const b200_tensor *dx, b200_stream s);
/* layer_norm: dx for the input, plus per-CTA partial rows of dgamma = sum(dy*xhat)
 * and dbeta = sum(dy) as [n_partials, d_model] tensors the caller column-sums with
 * b200_launch_reduce (deterministic; no atomics).  n_partials comes from
 * b200_layer_norm_backward_partials.  gamma may be NULL. */
int32_t b200_layer_norm_backward_partials(const b200_tensor *input, int32_t *n_partials);
int32_t b200_launch_layer_norm_backward(const b200_tensor *input, const b200_tensor *dy,
                                        const b200_tensor *gamma, double eps,
                                        const b200_tensor *dx,
                                        const b200_tensor *partial_gamma,
                                        const b200_tensor *partial_beta, b200_stream s);
/* Same, plus an optional third partial `partial_dx` [n_partials, R]: per-CTA column sums of dx.  When x came from a
 * Linear (x = h·W + b [+ residual]), colsum(dx) IS that Linear's bias gradient (linear_bias_backward,
 * crates/burn-backend/src/backend/ops/modules/linear.rs:117-128): the [tokens, d] column reduce that would re-read
 * dx disappears.  NULL = not wanted. */
int32_t b200_launch_layer_norm_backward_ex(const b200_tensor *input, const b200_tensor *dy,
                                           const b200_tensor *gamma, double eps, const b200_tensor *dx,
                                           const b200_tensor *partial_gamma, const b200_tensor *partial_beta,
                                           const b200_tensor *partial_dx, b200_stream s);

/* Cross-entropy on logits, value and gradient in one row-resident pass:
 * picked[r] = log_softmax(logits[r])[targets[r]] (the tensor CrossEntropyLoss::forward_default
 * reduces with mean().neg(), crates/burn-nn/src/loss/cross_entropy.rs:171-197) and
 * dlogits = (softmax(logits) - onehot(targets)) * grad_scale, its gradient when the caller passes
 * grad_scale = upstream / N.  dlogits may alias logits.  targets i32 / i64 [N]. */
int32_t b200_launch_softmax_cross_entropy(const b200_tensor *logits, const b200_tensor *targets,
                                          double grad_scale, const b200_tensor *picked,
                                          const b200_tensor *dlogits, b200_stream s);

/* ------------------------------------------------ fused attention (forward) */
/* ModuleOps::attention (crates/burn-backend/src/backend/ops/modules/base.rs:822-830;
 * semantics: attention_fallback, ops/modules/attention.rs:15-90):
 *   out[B,H,Sq,Dv] = softmax(q·kᵀ·scale, masked positions = mask_value) · v
 * mask: optional bool [B|1, H|1, Sq, Sk], nonzero = masked; is_causal additionally
 * masks col > row + (Sk - Sq) (fully masked KV blocks are skipped).  mask_value is
 * -inf for ModuleOps::attention, -1e9 for burn-nn's MultiHeadAttention
 * (mha.rs:291-296).  weights (optional, [B,H,Sq,Sk]) receives the softmax output —
 * what MultiHeadAttention keeps for its backward.  Head dim 64, f32 storage, tf32
 * tensor-core products; other head dims return B200_ERR_UNSUPPORTED (use the chain). */
int32_t b200_launch_attention(const b200_tensor *q, const b200_tensor *k,
                              const b200_tensor *v, const b200_tensor *mask,
                              double scale, double mask_value, int32_t is_causal,
                              const b200_tensor *out, const b200_tensor *weights,
                              b200_stream s);

/* Backward of the attention core given the saved weights P and context `out`
 * (what burn-autodiff's reverse walk over mha.rs:253-311 computes), first half in one
 * kernel: dP = d_out·vᵀ, dS = P ∘ (dP − rowsum(d_out ∘ out)) · scale, dq = dS·k.  `ds`
 * [B,H,Sq,Sk] is written for the two remaining products, plain float_matmul calls:
 * dk = dSᵀ·q and dv = Pᵀ·d_out.  Masked positions need no mask (P = 0 there). */
int32_t b200_launch_attention_backward(const b200_tensor *d_out, const b200_tensor *k,
                                       const b200_tensor *v, const b200_tensor *out,
                                       const b200_tensor *weights, double scale,
                                       int32_t is_causal, const b200_tensor *dq,
                                       const b200_tensor *ds, b200_stream s);

/* Flash-style attention for training (burn_b200/csrc/attention_flash.cu): the same ModuleOps::attention
 * forward, but instead of the weights it saves per-row softmax statistics — `stats` f32 [B,H,Sq,4] =
 * (row max in the base-2 domain, 1 / row sum, delta (filled by the backward), unused) — so no
 * [B,H,Sq,Sk] tensor is ever written.  mask: optional bool [B|1,H|1,Sq,Sk]; is_causal as above. */
int32_t b200_launch_attention_flash(const b200_tensor *q, const b200_tensor *k,
                                    const b200_tensor *v, const b200_tensor *mask,
                                    double scale, double mask_value, int32_t is_causal,
                                    const b200_tensor *out, const b200_tensor *stats,
                                    b200_stream s);
/* Its backward (what burn-autodiff's reverse walk over attention.rs:15-90 / mha.rs:253-311 computes:
 * matmul backward ×2, softmax backward, mask_fill backward, scaling): recomputes the weights tile by
 * tile from q, k and `stats`; dq in one kernel (query-row CTAs), dk and dv in a second (key-row CTAs),
 * all three accumulated in TMEM, no atomics (bit-reproducible).  Pass the forward's mask / scale /
 * mask_value / is_causal.  `stats` is updated in place (delta). */
int32_t b200_launch_attention_flash_backward(const b200_tensor *d_out, const b200_tensor *q,
                                             const b200_tensor *k, const b200_tensor *v,
                                             const b200_tensor *out, const b200_tensor *stats,
                                             const b200_tensor *mask, double scale,
                                             double mask_value, int32_t is_causal,
                                             const b200_tensor *dq, const b200_tensor *dk,
                                             const b200_tensor *dv, b200_stream s);

/* ------------------------------------------------ optimizer */
/* Multi-tensor Adam over one flat buffer, in place: the op sequence of
 * AdaptiveMomentum::transform + Adam::step
 * (crates/burn-optim/src/optim/adam.rs:149-210, :80-84) in one pass.
 * coef = device [2] f32 {sqrt(1-b2^t)/(1-b1^t), eps*sqrt(1-b2^t)}: the
 * time-dependent scalars live in device memory so a captured step replays. */
int32_t b200_launch_adam(const b200_tensor *param, const b200_tensor *moment1,
                         const b200_tensor *moment2, const b200_tensor *grad,
                         const b200_tensor *coef, double lr, double beta1,
                         double beta2, b200_stream s);

/* ------------------------------------------------ collectives */
/* DistributedOps::{all_reduce, sync_collective}
 * (crates/burn-backend/src/backend/distributed/ops.rs:116-131; reference impl
 * crates/burn-cubecl/src/ops/distributed.rs:17-55).  One communicator per
 * process (one process per GPU); the 128-byte unique id is produced on rank 0
 * and distributed by the host (torch.distributed / any bootstrap). */
#define B200_NCCL_UNIQUE_ID_BYTES 128
typedef void *b200_comm;
typedef enum { B200_REDUCE_SUM = 0, B200_REDUCE_MEAN = 1 } b200_reduce_op;
int32_t b200_comm_unique_id(uint8_t id[B200_NCCL_UNIQUE_ID_BYTES]);
int32_t b200_comm_init(b200_comm *out, const uint8_t id[B200_NCCL_UNIQUE_ID_BYTES],
                       int32_t rank, int32_t world_size);
int32_t b200_comm_destroy(b200_comm comm);
/* In-place all-reduce (sendbuff == recvbuff, as the reference does) on the
 * communicator's dedicated stream, ordered after everything already queued on
 * `producer` (event fence). */
int32_t b200_all_reduce(b200_comm comm, void *ptr, uint64_t count, int32_t dtype,
                        int32_t op, b200_stream producer);
/* Bucketed variant: n tensors reduced inside one ncclGroup. */
int32_t b200_all_reduce_multi(b200_comm comm, void *const *ptrs,
                              const uint64_t *counts, int32_t n, int32_t dtype,
                              int32_t op, b200_stream producer);
/* Makes `consumer` wait for every collective issued so far
 * (DistributedOps::sync_collective). */
int32_t b200_collective_sync(b200_comm comm, b200_stream consumer);
/* Finer-grained fences: record `e` on the collective stream behind the collectives issued so far,
 * and make a stream wait for an event — so the optimizer launches of one gradient bucket wait for
 * that bucket's all-reduce only, while later buckets are still in flight. */
int32_t b200_collective_mark(b200_comm comm, b200_event e);
int32_t b200_stream_wait_event(b200_stream s, b200_event e);

/* Single-process multi-device form — the reference's DDP contract: one process, a thread per GPU, ONE sync thread
 * issuing all_reduce for every device (crates/burn-backend/src/backend/distributed/server.rs:100-139).
 * b200_comm_init_all = ncclCommInitAll over `devices`; b200_all_reduce_group issues the n per-device all-reduces of
 * one tensor inside a single ncclGroupStart/End, so one thread can drive all communicators without deadlock.
 * ptrs[i] lives on devices[i]; producers[i] is a cudaStream_t of that device (NULL = its default stream). */
int32_t b200_comm_init_all(b200_comm *out, const int32_t *devices, int32_t n);
int32_t b200_all_reduce_group(const b200_comm *comms, void *const *ptrs, uint64_t count, int32_t n,
                              int32_t dtype, int32_t op, void *const *producers);
int32_t b200_comm_host_sync(b200_comm comm); /* block the host until the communicator's stream is idle */

/* ------------------------------------------------ peer-memory collectives (NVLink loads / stores, no NCCL) */
/* Every rank owns one REGION (b200_peer_alloc: cudaMalloc'd, zeroed, IPC-exportable) whose first
 * b200_peer_flag_bytes() bytes are synchronisation flags and whose remainder — the data area, b200_peer_data —
 * holds the flat gradient and parameter buckets at the SAME offsets on every rank.  A group maps all regions into
 * the caller: between processes through the 64-byte handles of b200_peer_export (gathered by the host bootstrap),
 * inside one process through b200_peer_group_create_local (peer access between the listed devices; out[i] is
 * device i's group and owns its region).
 *
 * b200_launch_peer_all_reduce: DistributedOps::all_reduce on `count` f32 at element `offset` of the data area —
 *   reduce-scatter + all-gather in one kernel, each rank reducing 1/N and storing it into every rank's bucket.
 * b200_launch_peer_adam: the gradient all-reduce (Mean) FUSED with the Adam step that consumes it
 *   (crates/burn-optim/src/optim/adam.rs:149-210): each rank reduces 1/N of the bucket, updates that shard of the
 *   parameters with its shard of the moments (optimizer state sharded N ways; `moment1/2` are full-size local
 *   buffers of which only the owned shard is used) and stores the new parameters into every rank's bucket.
 * Both run on the group's own high-priority stream, ordered after `producer` (event fence), identically on every
 * rank and in the same order; `slot` (one per bucket) names the flag words.  Results are bit-identical on all ranks.
 * b200_peer_sync = sync_collective: `consumer` waits for everything issued so far. */
#define B200_PEER_MAX_RANKS 8
#define B200_PEER_HANDLE_BYTES 64
typedef void *b200_peer_group;
int32_t b200_peer_alloc(void **out, uint64_t bytes);
int32_t b200_peer_free(void *ptr);
int32_t b200_peer_export(void *ptr, uint8_t handle[B200_PEER_HANDLE_BYTES]);
uint64_t b200_peer_flag_bytes(void);
int32_t b200_peer_group_create(b200_peer_group *out, int32_t rank, int32_t world, void *local_region,
                               uint64_t bytes, const uint8_t *handles /* world x 64, own entry ignored */);
int32_t b200_peer_group_create_local(b200_peer_group *out /* [n] */, const int32_t *devices, int32_t n, uint64_t bytes);
void *b200_peer_data(b200_peer_group group);
int32_t b200_peer_group_destroy(b200_peer_group group);
int32_t b200_launch_peer_all_reduce(b200_peer_group group, uint64_t offset, uint64_t count, int32_t op,
                                    int32_t slot, b200_stream producer);
int32_t b200_launch_peer_adam(b200_peer_group group, uint64_t grad_offset, uint64_t param_offset,
                              void *moment1, void *moment2, const void *coef, uint64_t count,
                              double lr, double beta1, double beta2, int32_t slot, b200_stream producer);
int32_t b200_peer_sync(b200_peer_group group, b200_stream consumer);
int32_t b200_peer_mark(b200_peer_group group, b200_event e);
int32_t b200_peer_host_sync(b200_peer_group group);

/* NVRTC specialisation self-test: generates and compiles (sm_100a, no device needed, nothing loaded)
 * the specialised elementwise and fuse-on-read reduce kernels of the bench chain; reports the cubin bytes. */
int32_t b200_jit_selftest(uint64_t *cubin_bytes_total);
/* The specialised-kernel cache is a bounded LRU (B200_JIT_CACHE_MAX, default 512 kernels; an evicted cubin is
 * unloaded): how many kernels it holds and how many it has evicted. */
int32_t b200_jit_cache_stats(uint64_t *entries, uint64_t *evictions);

/* ------------------------------------------------ introspection */
/* Number of kernels this library has launched since load (bench.py's
 * gpu_launches claim) and a reset. */
uint64_t b200_launch_count(void);
void b200_launch_count_reset(void);

#ifdef __cplusplus
}
#endif
#endif /* BURN_B200_H */
