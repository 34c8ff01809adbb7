// tcgen05 TF32 GLU GEMM — placeholder until the tensor-core kernel lands (returns "unsupported").
#include "common.cuh"
#include "internal.cuh"
namespace sg {
int glu_gemm_tc(int, int, int, const float*, int, const float*, const float*, const float*, const float*,
                float*, int, float*, float*, int, cudaStream_t) {
  return -1;
}
}  // namespace sg
