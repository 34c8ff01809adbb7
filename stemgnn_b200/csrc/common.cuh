// Shared host/device helpers for the stemgnn_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/stemgnn_b200.h"

namespace sg {

// ---- error plumbing: C ABI returns int, message is thread-local -------------------------
void set_error(const char* fmt, ...);
void clear_error();
void count_launch();                       // bumps the counter behind stemgnn_launch_count()
// optional event pair recorded around the GRU recurrence kernel (stemgnn_profile_gru)
struct ProfileHook { cudaEvent_t start, stop; };
ProfileHook* profile_hook();

#define SG_CHECK(cond, ...)                      \
  do {                                           \
    if (!(cond)) {                               \
      sg::set_error(__VA_ARGS__);                \
      return 1;                                  \
    }                                            \
  } while (0)

#define SG_CUDA(expr)                                                                  \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      sg::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                    __LINE__);                                                         \
      return 2;                                                                        \
    }                                                                                  \
  } while (0)

#define SG_LAUNCH_CHECK(name)                                                       \
  do {                                                                              \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess) {                                                        \
      sg::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));       \
      return 3;                                                                     \
    }                                                                               \
    sg::count_launch();                                                             \
  } while (0)

#define SG_TRY(expr)            \
  do {                          \
    int _rc = (expr);           \
    if (_rc != 0) return _rc;   \
  } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int pad4(int n) { return (n + 3) & ~3; }

// ---- device helpers ---------------------------------------------------------------------
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float leaky_(float x, float a) { return x >= 0.f ? x : x * a; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum for blockDim.x <= 1024 (scratch: 32 floats of shared memory).
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = (lane < nw) ? scratch[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

// Philox4x32-10 counter RNG: the dropout keep-mask of attention element (b,i,j) is a pure
// function of (seed, offset, linear index) so backward regenerates it instead of storing it.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}
// keep-probability test for linear element index `idx` (one 32-bit lane of a Philox block)
__device__ __forceinline__ bool dropout_keep(uint64_t seed, uint64_t offset, uint64_t idx, float p) {
  const uint64_t blk = (idx >> 2) + offset;
  uint4 c = make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), 0u, 0u);
  uint4 r = philox4x32_10(c, make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  const uint32_t lane = (uint32_t)(idx & 3);
  const uint32_t v = lane == 0 ? r.x : lane == 1 ? r.y : lane == 2 ? r.z : r.w;
  // uniform in [0,1): keep iff u >= p  (matches "mask = rand >= p" in the tests)
  return (float)(v >> 8) * (1.0f / 16777216.0f) >= p;
}

}  // namespace sg
