// Latent-correlation graph construction after the GRU  (reference: models/base_model.py)
//   self_graph_attention  :151-162   data[b,i,j] = key[b,i] + query[b,j] -> LeakyReLU -> softmax_j
//                                    -> Dropout
//   latent_correlation    :140-147   mean over batch, degree (pre-symmetrise), symmetrise,
//                                    L = D^(D - A)D^ with D^ = diag(1/(sqrt(deg)+1e-7))
// The (B,N,N) attention tensor is never materialised: row maxima follow from monotonicity of
// LeakyReLU (max_j lrelu(k_i+q_j) = lrelu(k_i + max_j q_j)), the softmax denominators take one
// pass, and the batch mean is accumulated on the fly (2 B N^2 exps in total).
#include "common.cuh"
#include "internal.cuh"

namespace sg {

// x (B,W,N) -> xs (N,B,W) [GRU sequence layout] and x_bnw (B,N,W) [spectral-block layout]
__global__ void prep_layouts_kernel(const float* __restrict__ x, float* __restrict__ xs,
                                    float* __restrict__ x_bnw, int B, int W, int N, float* __restrict__ x_pad,
                                    int ld_pad) {
  const long long total = (long long)B * W * N;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(idx % N);
    const int t = (int)((idx / N) % W);
    const int b = (int)(idx / ((long long)N * W));
    const float v = x[idx];
    xs[((long long)n * B + b) * W + t] = v;
    x_bnw[((long long)b * N + n) * W + t] = v;
    if (x_pad != nullptr) x_pad[((long long)b * W + t) * ld_pad + n] = v;
  }
}

// one CTA per attention row i
__global__ void __launch_bounds__(256) attention_mean_kernel(AttnArgs a) {
  extern __shared__ float sm[];
  const int B = a.B, N = a.N, i = blockIdx.x;
  float* s_key = sm;            // [B]
  float* s_m = sm + B;          // [B]
  float* s_zinv = sm + 2 * B;   // [B]
  __shared__ float red[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;

  for (int b = threadIdx.x; b < B; b += blockDim.x) s_key[b] = a.key[(long long)b * N + i];
  __syncthreads();
  // phase 1: one warp per batch element: softmax denominator of row (b,i)
  for (int b = wid; b < B; b += nw) {
    const float ki = s_key[b];
    // row maximum without materialising the row: max_j lrelu(k_i + q_j) = lrelu(k_i + max_j q_j) (file header); the query row
    // is read twice from L1/L2 by the same warp instead of once more by a separate kernel launch
    float qm = -INFINITY;
    for (int j = lane; j < N; j += 32) qm = fmaxf(qm, a.query[(long long)b * N + j]);
    qm = warp_max(qm);
    const float m = leaky_(ki + qm, a.alpha);
    float z = 0.f;
    for (int j = lane; j < N; j += 32)
      z += expf(leaky_(ki + a.query[(long long)b * N + j], a.alpha) - m);
    z = warp_sum(z);
    if (lane == 0) {
      s_m[b] = m;
      s_zinv[b] = 1.0f / z;
      if (a.row_m != nullptr) {
        a.row_m[(long long)b * N + i] = m;
        a.row_zinv[(long long)b * N + i] = 1.0f / z;
      }
    }
  }
  __syncthreads();
  // phase 2: thread per column j: batch mean of the (dropped) probabilities
  const float scale = (a.use_dropout ? 1.0f / (1.0f - a.p) : 1.0f) / (float)B;
  const uint64_t doff = a.offset + (a.offset_dev != nullptr ? (uint64_t)__ldg(a.offset_dev) : 0ull);
  float dsum = 0.f;
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    float acc = 0.f;
    for (int b = 0; b < B; ++b) {
      bool keep = true;
      if (a.use_dropout) {
        const uint64_t lin = ((uint64_t)b * N + i) * N + j;
        keep = a.mask != nullptr ? (a.mask[lin] != 0) : dropout_keep(a.seed, doff, lin, a.p);
      }
      if (keep)
        acc += expf(leaky_(s_key[b] + a.query[(long long)b * N + j], a.alpha) - s_m[b]) * s_zinv[b];
    }
    acc *= scale;
    a.a_raw[(long long)i * N + j] = acc;
    dsum += acc;
  }
  dsum = block_sum(dsum, red);
  if (threadIdx.x == 0) a.deg[i] = dsum;
}

// attention_sym = (A + A^T)/2 ;  L = D^ (diag(deg) - attention_sym) D^ ;  mul_L[0] = 0, mul_L[1] = L
__global__ void laplacian_kernel(const float* __restrict__ a_raw, const float* __restrict__ deg,
                                 float* __restrict__ attention, float* __restrict__ mul_L, int N,
                                 float* __restrict__ L_pad, int ld_pad) {
  const long long total = (long long)N * N;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / N), j = (int)(idx % N);
    const float asym = 0.5f * (a_raw[idx] + a_raw[(long long)j * N + i]);
    const float di = 1.0f / (sqrtf(deg[i]) + 1e-7f);
    const float dj = 1.0f / (sqrtf(deg[j]) + 1e-7f);
    const float inner = ((i == j) ? deg[i] : 0.f) - asym;
    attention[idx] = asym;
    mul_L[idx] = 0.f;
    const float l = di * (inner * dj);
    mul_L[total + idx] = l;
    if (L_pad != nullptr) L_pad[(long long)(3 * i) * ld_pad + j] = l;    // row i*3 + k', k' = 0
  }
}

int launch_prep_layouts(const float* x, float* xs, float* x_bnw, int B, int W, int N, cudaStream_t st, float* x_pad,
                        int ld_pad) {
  const long long total = (long long)B * W * N;
  const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
  prep_layouts_kernel<<<blocks, 256, 0, st>>>(x, xs, x_bnw, B, W, N, x_pad, ld_pad);
  SG_LAUNCH_CHECK("prep_layouts_kernel");
  return 0;
}

int launch_attention(const AttnArgs& a, float* qmax, cudaStream_t st) {
  (void)qmax;      // (round 1 launched a separate row-max kernel into this buffer; the attention kernel computes it now)
  const size_t smem = (size_t)3 * a.B * sizeof(float);
  SG_CHECK(smem <= 40 * 1024, "attention: batch %d too large for the row kernel", a.B);
  attention_mean_kernel<<<a.N, 256, smem, st>>>(a);
  SG_LAUNCH_CHECK("attention_mean_kernel");
  return 0;
}

int launch_laplacian(const float* a_raw, const float* deg, float* attention, float* mul_L, int N,
                     cudaStream_t st, float* L_pad, int ld_pad) {
  const long long total = (long long)N * N;
  const int blocks = (int)((total + 255) / 256 < 2048 ? (total + 255) / 256 : 2048);
  laplacian_kernel<<<blocks, 256, 0, st>>>(a_raw, deg, attention, mul_L, N, L_pad, ld_pad);
  SG_LAUNCH_CHECK("laplacian_kernel");
  return 0;
}

}  // namespace sg
