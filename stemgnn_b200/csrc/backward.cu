// Backward of the hot path — placeholder (filled in below the forward milestone).
#include "common.cuh"
#include "internal.cuh"
namespace sg {
int model_backward(const stemgnn_dims_t*, const stemgnn_params_t*, const stemgnn_fwd_opts_t*,
                   const float*, const float*, const float*, const stemgnn_grads_t*, float*, void*,
                   size_t, cudaStream_t) {
  set_error("stemgnn_model_backward: not implemented yet");
  return 1;
}
}  // namespace sg
