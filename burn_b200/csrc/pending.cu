// Entry points declared in burn_b200.h whose kernels are not written yet.
// They fail loudly (no fallback).
#include "common.cuh"
using namespace b200;

extern "C" int32_t b200_launch_softmax(const b200_tensor *, const b200_tensor *, int32_t, b200_stream) {
  return fail(B200_ERR_UNSUPPORTED, "b200_launch_softmax is not implemented yet");
}
extern "C" int32_t b200_launch_layer_norm(const b200_tensor *, const b200_tensor *, const b200_tensor *, double,
                                          const b200_tensor *, b200_stream) {
  return fail(B200_ERR_UNSUPPORTED, "b200_launch_layer_norm is not implemented yet");
}
