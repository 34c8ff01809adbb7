// Train-step tail on the device (reference: models/handler.py:160-166 — MSELoss(mean), backward,
// RMSprop / Adam step, and the per-step `float(loss)` host sync that this removes).
//   * stemgnn_mse_loss_grad : d_forecast = 2 (f - y) / n and loss_accum += mean((f - y)^2), one block, fixed order;
//   * stemgnn_optimizer_step: ONE launch over the flat parameter / gradient / state buffers (torch.optim.RMSprop and
//     torch.optim.Adam update rules with their default flags), learning rate and step count read from device memory so
//     that a captured CUDA graph can be replayed while the schedule changes;
//   * stemgnn_counters_tick : bumps the device-side step counter and the Philox dropout offset after a step.
#include "common.cuh"
#include "internal.cuh"

namespace sg {
namespace {

__global__ void __launch_bounds__(1024) mse_loss_grad_kernel(const float* __restrict__ f, const float* __restrict__ y,
                                                             long long n, float* __restrict__ d_f,
                                                             float* __restrict__ loss_accum) {
  __shared__ float red[32];
  const float scale = 2.0f / (float)n;
  float acc = 0.f;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const float e = f[i] - y[i];
    d_f[i] = scale * e;
    acc = fmaf(e, e, acc);
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) loss_accum[0] += acc / (float)n;     // single writer: deterministic
}

// kind 0: RMSprop (alpha = h0, eps), kind 1: Adam (beta1 = h0, beta2 = h1, eps, bias correction with step + 1)
__global__ void __launch_bounds__(256) optimizer_step_kernel(int kind, float* __restrict__ p, const float* __restrict__ g,
                                                             float* __restrict__ s1, float* __restrict__ s2, long long n,
                                                             const float* __restrict__ lr_dev, float h0, float h1, float eps,
                                                             const unsigned long long* __restrict__ step) {
  const float lr = __ldg(lr_dev);
  float bc1 = 1.f, bc2s = 1.f;
  if (kind == 1) {
    const float t = (float)(__ldg(step) + 1ull);
    bc1 = 1.f - powf(h0, t);
    bc2s = sqrtf(1.f - powf(h1, t));
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i];
    if (kind == 0) {
      const float sq = h0 * s1[i] + (1.f - h0) * gi * gi;          // square_avg.mul_(alpha).addcmul_(g, g, 1 - alpha)
      s1[i] = sq;
      p[i] -= lr * gi / (sqrtf(sq) + eps);                          // p.addcdiv_(g, sqrt(square_avg) + eps, -lr)
    } else {
      const float m = h0 * s1[i] + (1.f - h0) * gi;
      const float v = h1 * s2[i] + (1.f - h1) * gi * gi;
      s1[i] = m;
      s2[i] = v;
      p[i] -= (lr / bc1) * m / (sqrtf(v) / bc2s + eps);
    }
  }
}

__global__ void counters_tick_kernel(unsigned long long* step, unsigned long long* dropout_ctr, unsigned long long inc) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (step != nullptr) step[0] += 1ull;
    if (dropout_ctr != nullptr) dropout_ctr[0] += inc;
  }
}

}  // namespace
}  // namespace sg

using namespace sg;

extern "C" {

int stemgnn_mse_loss_grad(const float* forecast, const float* target, long long n, float* d_forecast, float* loss_accum,
                          stemgnn_stream_t stream) {
  clear_error();
  SG_CHECK(forecast && target && d_forecast && loss_accum && n > 0, "mse_loss_grad: bad arguments");
  mse_loss_grad_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(forecast, target, n, d_forecast, loss_accum);
  SG_LAUNCH_CHECK("mse_loss_grad_kernel");
  return 0;
}

int stemgnn_optimizer_step(int kind, float* params, const float* grads, float* state1, float* state2, long long n,
                           const float* lr_dev, float h0, float h1, float eps, const unsigned long long* step_dev,
                           stemgnn_stream_t stream) {
  clear_error();
  SG_CHECK((kind == 0 || kind == 1) && params && grads && state1 && lr_dev && n > 0, "optimizer_step: bad arguments");
  SG_CHECK(kind == 0 || (state2 && step_dev), "optimizer_step: Adam needs the second moment buffer and the step counter");
  const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  optimizer_step_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(kind, params, grads, state1, state2, n,
                                                                               lr_dev, h0, h1, eps, step_dev);
  SG_LAUNCH_CHECK("optimizer_step_kernel");
  return 0;
}

int stemgnn_counters_tick(unsigned long long* step_dev, unsigned long long* dropout_counter_dev, unsigned long long dropout_inc,
                          stemgnn_stream_t stream) {
  clear_error();
  counters_tick_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(step_dev, dropout_counter_dev, dropout_inc);
  SG_LAUNCH_CHECK("counters_tick_kernel");
  return 0;
}

}  // extern "C"
