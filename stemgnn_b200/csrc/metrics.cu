// Validation metrics on the device (reference: models/handler.py:74-82 `validate`,
// data_loader/forecast_dataloader.py:25-38 `de_normalized`, utils/math_utils.py:24-74 `MAPE/MAE/RMSE/evaluate`).
//
// The reference copies every batch's forecast and target to the host and runs numpy reductions there (158 ms per
// validation on ECG, 583 ms on PeMS07 — SURVEY.md §6).  Here the (count, H, N) forecast / target tensors stay in HBM;
// one pass de-normalises (float64, the reference's dtype) and accumulates per node the six sums
//   raw:  sum min(|e|/|y| + 1e-5, 5), sum |e|, sum e^2      norm: the same on the normalised values
// into per-chunk partials, and a second kernel adds the partials in a fixed order (deterministic results).
// The host receives 6 x N doubles (the per-node vectors; the overall scores are their means).
#include "common.cuh"
#include "internal.cuh"

namespace sg {

namespace {

// numpy semantics of np.minimum(ratio, 5): NaN (0/0) propagates
__device__ __forceinline__ double clip_ape(double ratio) { return ratio != ratio ? ratio : fmin(ratio, 5.0); }

__global__ void __launch_bounds__(128) metrics_partial_kernel(const double* __restrict__ forecast,
                                                              const float* __restrict__ target, long long rows, int N,
                                                              int method, const double* __restrict__ scale,
                                                              const double* __restrict__ shift, int rows_per_chunk,
                                                              double* __restrict__ partial) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const long long r0 = (long long)blockIdx.y * rows_per_chunk;
  const long long r1 = r0 + rows_per_chunk < rows ? r0 + rows_per_chunk : rows;
  const double sc = method ? scale[n] : 1.0, sh = method ? shift[n] : 0.0;
  double a[6] = {0, 0, 0, 0, 0, 0};
  for (long long r = r0; r < r1; ++r) {
    const double fn = forecast[r * N + n];
    const double tn = (double)target[r * N + n];
    const double f = method ? fn * sc + sh : fn;       // de_normalized(): data * std + mean | data * span + min
    const double t = method ? tn * sc + sh : tn;
    const double e = f - t, en = fn - tn;
    a[0] += clip_ape(fabs(e) / fabs(t) + 1e-5);
    a[1] += fabs(e);
    a[2] += e * e;
    a[3] += clip_ape(fabs(en) / fabs(tn) + 1e-5);
    a[4] += fabs(en);
    a[5] += en * en;
  }
  double* out = partial + (size_t)blockIdx.y * 6 * N;
#pragma unroll
  for (int k = 0; k < 6; ++k) out[(size_t)k * N + n] = a[k];
}

__global__ void metrics_reduce_kernel(const double* __restrict__ partial, int chunks, int N, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 6 * N) return;
  double acc = 0.0;
  for (int c = 0; c < chunks; ++c) acc += partial[(size_t)c * 6 * N + i];   // fixed order
  out[i] = acc;
}

}  // namespace

int eval_metrics(const double* forecast, const float* target, long long count, int H, int N, int method,
                 const double* scale, const double* shift, double* partial, int chunks, double* out, cudaStream_t st) {
  const long long rows = count * H;
  SG_CHECK(rows > 0 && N > 0 && chunks > 0, "eval_metrics: bad sizes");
  const int rpc = (int)((rows + chunks - 1) / chunks);
  dim3 grid(ceil_div(N, 128), chunks);
  metrics_partial_kernel<<<grid, 128, 0, st>>>(forecast, target, rows, N, method, scale, shift, rpc, partial);
  SG_LAUNCH_CHECK("metrics_partial_kernel");
  metrics_reduce_kernel<<<ceil_div(6 * N, 256), 256, 0, st>>>(partial, chunks, N, out);
  SG_LAUNCH_CHECK("metrics_reduce_kernel");
  return 0;
}

}  // namespace sg

using namespace sg;

extern "C" int stemgnn_eval_metrics(const double* forecast_norm, const float* target_norm, long long count, int H, int N,
                                    int method, const double* scale, const double* shift, double* partial,
                                    int chunks, double* sums, stemgnn_stream_t stream) {
  clear_error();
  SG_CHECK(forecast_norm && target_norm && partial && sums, "eval_metrics: null argument");
  SG_CHECK(method == 0 || (scale && shift), "eval_metrics: normalisation statistics missing");
  return eval_metrics(forecast_norm, target_norm, count, H, N, method, scale, shift, partial, chunks, sums,
                      static_cast<cudaStream_t>(stream));
}
