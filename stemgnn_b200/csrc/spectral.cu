// Spectral block helpers  (reference: models/base_model.py StockBlockLayer :16-75)
//
// The two Fourier transforms of spe_seq_cell are LINEAR maps applied next to Linear layers, so they
// are folded into the adjacent weights once per weight version instead of being executed per row:
//   * rfft(W, two-sided) o GLU-0 Linear (base_model.py:49-53)  ->  time-domain weights (d x 3W)
//     (the k=0 Chebyshev channel is identically zero, base_model.py:129, so its W columns vanish);
//   * irfft(T) o sum_k (.)@weight[k] o {forecast, backcast} Linear (:58,:65-71)
//                                               ->  one (8T x (T+W)) map on [real3 | imag3];
//     irfft(onesided=False) reads bins 0..T/2 only and ignores Im(DC), Im(Nyquist),
//     which shows up here as exactly-zero rows of the folded matrix.
// What remains per row are plain GEMMs (gemm.cuh / glu_tc.cu) plus the small head kernels below.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "internal.cuh"
#include "gemm.cuh"

namespace sg {

// ---- fold: DFT_W into the first GLU layer ------------------------------------------------------
// w_in: (d, 4W) Linear weight acting on [k*W+f] real (chain 0) or imag (chain 1) spectra;
// w_out: (d, 3W) acting on the time samples g[k'*W+t], k' = k-1 in 0..2.
//   real_f =  sum_t g_t cos(2 pi f t / W)      imag_f = -sum_t g_t sin(2 pi f t / W)
// kfirst/nk: Chebyshev channels kept (model path: 1,3 — channel 0 is zero; stage API: 0,4).
__global__ void fold_in_kernel(const float* __restrict__ w_in, float* __restrict__ w_out, int d, int W,
                               int chain, int kfirst, int nk) {
  const int o = blockIdx.x;
  for (int c = threadIdx.x; c < nk * W; c += blockDim.x) {
    const int kp = c / W, t = c % W;
    const float* wrow = w_in + (long long)o * 4 * W + (kp + kfirst) * W;
    float acc = 0.f;
    for (int f = 0; f < W; ++f) {
      const int ph = (f * t) % W;
      const float ang = 2.0f * (float)ph / (float)W;
      const float tw = chain == 0 ? cospif(ang) : -sinpif(ang);
      acc = fmaf(wrow[f], tw, acc);
    }
    w_out[(long long)o * nk * W + c] = acc;
  }
}

// ---- fold: inverse real DFT table -------------------------------------------------------------
// ic[chain][f][t]: contribution of Re (chain 0) / Im (chain 1) of bin f to output sample t of
// irfft(n=T) (norm 1/T, bins 0..T/2, Im of DC / Nyquist ignored).
__global__ void irfft_table_kernel(float* __restrict__ ic, int T) {
  const int total = 2 * T * T;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int chain = idx / (T * T), f = (idx / T) % T, t = idx % T;
    const int half = T / 2;
    float v = 0.f;
    if (f <= half) {
      const bool edge = (f == 0) || ((T % 2 == 0) && f == half);
      const int ph = (f * t) % T;
      const float ang = 2.0f * (float)ph / (float)T;
      if (chain == 0) v = (edge ? 1.0f : 2.0f) * cospif(ang) / (float)T;
      else v = edge ? 0.f : -2.0f * sinpif(ang) / (float)T;
    }
    ic[idx] = v;
  }
}

// ---- GFT scatter epilogue ----------------------------------------------------------------------
// GEMM rows m = k'*N + n (k' = 0..2 -> Chebyshev terms 1..3), cols c = b*W + t.
// Destination: G[(b*N + n) * 3W + k'*W + t]  (row-major (B*N) x 3W activation for the GLU chain).
struct EpiGftScatter {
  float* G; int N, W; int atomic;    // atomic = 1: split-K partial sums are added into a zeroed G
  __device__ __forceinline__ void store4(int, int m, int n, int valid, float4 v, float4) const {
    const int kp = m / N, node = m - kp * N;
    const float vals[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < valid) {
        const int c = n + j, b = c / W, t = c - b * W;
        float* dst = G + ((long long)b * N + node) * (3 * W) + kp * W + t;
        if (atomic) atomicAdd(dst, vals[j]);
        else *dst = vals[j];
      }
    }
  }
};

// G[(b*N + n)*3W + k'*W + t] = sum_z P[z][(k'*N + n)*(B*W) + b*W + t]
__global__ void gft_reduce_scatter_kernel(const float* __restrict__ P, int ks, float* __restrict__ G, int B, int N,
                                          int W) {
  const long long total = (long long)B * N * 3 * W;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int t = (int)(idx % W), kp = (int)((idx / W) % 3);
  const long long bn = idx / (3 * W);
  const int n = (int)(bn % N), b = (int)(bn / N);
  const long long src = ((long long)kp * N + n) * ((long long)B * W) + (long long)b * W + t;
  const long long stride = (long long)3 * N * B * W;
  float acc = 0.f;
  for (int z = 0; z < ks; ++z) acc += P[(long long)z * stride + src];
  G[idx] = acc;
}

// same reduction, additionally emitting the 16-bit operand images of G the kind::f16 GLU chain reads (row pitch `ldh`
// halves, zero padded): mode 0 = fp16 hi/lo split, 1 = bf16 (hi only).  One thread per image element.
__global__ void gft_reduce_scatter_h_kernel(const float* __restrict__ P, int ks, float* __restrict__ G, int B, int N, int W,
                                            unsigned short* __restrict__ g_hi, unsigned short* __restrict__ g_lo, int ldh,
                                            int bf16) {
  const long long total = (long long)B * N * ldh;
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % ldh);
  const long long bn = idx / ldh;
  unsigned short hi = 0, lo = 0;
  if (c < 3 * W) {
    const int t = c % W, kp = c / W;
    const int n = (int)(bn % N), b = (int)(bn / N);
    const long long src = ((long long)kp * N + n) * ((long long)B * W) + (long long)b * W + t;
    const long long stride = (long long)3 * N * B * W;
    float acc = 0.f;
    for (int z = 0; z < ks; ++z) acc += P[(long long)z * stride + src];
    G[bn * (3 * W) + c] = acc;
    if (bf16) {
      hi = __bfloat16_as_ushort(__float2bfloat16_rn(acc));
    } else {
      const float xs = fminf(fmaxf(acc, -65504.f), 65504.f);
      const __half h = __float2half_rn(xs);
      hi = __half_as_ushort(h);
      lo = __half_as_ushort(__float2half_rn(xs - __half2float(h)));
    }
  }
  g_hi[idx] = hi;
  if (!bf16) g_lo[idx] = lo;
}

// g_img != nullptr: also write the 16-bit images [hi (B*N x ldh) | lo] for the tensor-core chain; *g_ready tells the caller
int launch_gft(const float* mul_L, const float* x_bwn, float* G, float* skbuf, int B, int N, int W,
               cudaStream_t st, unsigned short* g_img, int ldh, int bf16, int* g_ready) {
  if (g_ready != nullptr) *g_ready = 0;
  // A = mul_L[1..3] viewed as (3N x N); B operand = x (B*W x N) read as B[n*ldb + k]
  // (3N x N) . (N x B*W) fills only ~54 CTAs: deterministic split-K (partials + fixed-order reduction)
  const int ks = skbuf != nullptr ? pick_ksplit(3 * N, B * W, N) : 1;
  if (ks > 1 && ks <= 8) {      // skbuf holds 8 partial products
    GemmOperands g = {mul_L + (long long)N * N, N, 0, x_bwn, N, 0, nullptr, 3 * N, B * W, N, ks};
    EpiPartial epi = {skbuf, B * W, (long long)3 * N * B * W};
    SG_TRY((launch_sgemm<false, true, false>(g, epi, 1, st, "gft_gemm_splitk")));
    if (g_img != nullptr) {
      const long long total = (long long)B * N * ldh;
      gft_reduce_scatter_h_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(skbuf, ks, G, B, N, W, g_img,
                                                                              g_img + (size_t)B * N * ldh, ldh, bf16);
      SG_LAUNCH_CHECK("gft_reduce_scatter_h_kernel");
      if (g_ready != nullptr) *g_ready = 1;
      return 0;
    }
    const long long total = (long long)B * N * 3 * W;
    gft_reduce_scatter_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(skbuf, ks, G, B, N, W);
    SG_LAUNCH_CHECK("gft_reduce_scatter_kernel");
    return 0;
  }
  GemmOperands g = {mul_L + (long long)N * N, N, 0, x_bwn, N, 0, nullptr, 3 * N, B * W, N};
  EpiGftScatter epi = {G, N, W, 0};
  return launch_sgemm<false, true, false>(g, epi, 1, st, "gft_gemm");
}

// ---- per-block head ----------------------------------------------------------------------------
// pre (R x ldp): cols [0,T) = igfted @ forecast.weight^T (no bias), cols [T,T+W) = igfted @ backcast.weight^T
//   forecast_source = sigmoid(pre_f + b_f)                         base_model.py:68
//   forecast        = forecast_source @ Wfr^T + b_fr               :69
//   backcast        = sigmoid(pre_b + b_b - (x @ Wsc^T + b_sc))    :70-72 (block 0 only)

// one warp per row: the T sigmoids are computed once (not once per output), staged in shared memory, then lanes
// o < W take the dot products against forecast_result.weight (transposed in shared memory: conflict-free)
__global__ void __launch_bounds__(256) block_head_kernel(HeadArgs a) {
  extern __shared__ float sm[];
  const int T = a.T, W = a.W;
  float* s_wfr = sm;                    // [T][W]   forecast_result.weight^T
  float* s_wsc = s_wfr + T * W;         // [W][W]   backcast_short_cut.weight^T
  float* s_fs = s_wsc + W * W;          // [8][T]
  float* s_x = s_fs + 8 * T;            // [8][W]
  for (int i = threadIdx.x; i < T * W; i += blockDim.x) s_wfr[(i % T) * W + i / T] = a.wfr[i];
  if (a.backcast_bnw != nullptr)
    for (int i = threadIdx.x; i < W * W; i += blockDim.x) s_wsc[(i % W) * W + i / W] = a.wsc[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* fs = s_fs + w * T;
  float* xr = s_x + w * W;
  for (int row = blockIdx.x * 8 + w; row < a.R; row += gridDim.x * 8) {
    const float* p = a.pre + (long long)row * a.ldp;
    for (int u = lane; u < T; u += 32) {
      const float f = sigmoidf_(p[u] + a.bf[u]);
      fs[u] = f;
      if (a.save_fs != nullptr) a.save_fs[(long long)row * T + u] = f;
    }
    if (a.backcast_bnw != nullptr)
      for (int t = lane; t < W; t += 32) xr[t] = a.x_bnw[(long long)row * W + t];
    __syncwarp();
    for (int o = lane; o < W; o += 32) {
      float acc = a.bfr[o];
      for (int u = 0; u < T; ++u) acc = fmaf(fs[u], s_wfr[u * W + o], acc);
      a.forecast[(long long)row * W + o] = acc;
      if (a.backcast_bnw != nullptr) {
        float sc = a.bsc[o];
        for (int t = 0; t < W; ++t) sc = fmaf(xr[t], s_wsc[t * W + o], sc);
        const float bc = sigmoidf_(p[T + o] + a.bb[o] - sc);
        a.backcast_bnw[(long long)row * W + o] = bc;
        const int b = row / a.N, n = row - b * a.N;
        a.backcast_bwn[((long long)b * W + o) * a.N + n] = bc;
      }
    }
    __syncwarp();
  }
}

int launch_block_head(const HeadArgs& a, cudaStream_t st) {
  const int blocks = a.R / 8 + 1 < 148 * 8 ? a.R / 8 + 1 : 148 * 8;
  const size_t smem = (size_t)(a.T * a.W + a.W * a.W + 8 * a.T + 8 * a.W) * sizeof(float);
  SG_CHECK(smem <= 48 * 1024, "block head: T=%d W=%d need %zu bytes of shared memory", a.T, a.W, smem);
  block_head_kernel<<<blocks, 256, smem, st>>>(a);
  SG_LAUNCH_CHECK("block_head_kernel");
  return 0;
}

// ---- model head:  fc(forecast_0 + forecast_1)   base_model.py:174-179, :97-101 ------------------
// out[b][h][n] = fc2(leaky_relu_0.01(fc0(f0 + f1)))[b][n][h]
__global__ void __launch_bounds__(128) model_head_kernel(const float* __restrict__ f0,
                                                         const float* __restrict__ f1,
                                                         const float* __restrict__ w0,
                                                         const float* __restrict__ b0,
                                                         const float* __restrict__ w2,
                                                         const float* __restrict__ b2,
                                                         float* __restrict__ out, int B, int N, int W,
                                                         int H) {
  extern __shared__ float sm[];
  float* s_w0 = sm;                 // [W][W]
  float* s_w2 = s_w0 + W * W;       // [H][W]
  float* s_b0 = s_w2 + H * W;       // [W]
  float* s_b2 = s_b0 + W;           // [H]
  for (int i = threadIdx.x; i < W * W; i += blockDim.x) s_w0[i] = w0[i];
  for (int i = threadIdx.x; i < H * W; i += blockDim.x) s_w2[i] = w2[i];
  for (int i = threadIdx.x; i < W; i += blockDim.x) s_b0[i] = b0[i];
  for (int i = threadIdx.x; i < H; i += blockDim.x) s_b2[i] = b2[i];
  __syncthreads();
  const long long total = (long long)B * N * H;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    // idx enumerates (b, h, n) so that the store is coalesced over n
    const int n = (int)(idx % N), h = (int)((idx / N) % H), b = (int)(idx / ((long long)N * H));
    const long long row = (long long)b * N + n;
    float acc = s_b2[h];
    for (int j = 0; j < W; ++j) {
      float hj = s_b0[j];
      for (int t = 0; t < W; ++t)
        hj = fmaf(f0[row * W + t] + f1[row * W + t], s_w0[j * W + t], hj);
      hj = leaky_(hj, 0.01f);
      acc = fmaf(hj, s_w2[h * W + j], acc);
    }
    out[idx] = acc;
  }
}

int launch_model_head(const float* f0, const float* f1, const float* w0, const float* b0,
                      const float* w2, const float* b2, float* out, int B, int N, int W, int H,
                      cudaStream_t st) {
  const long long total = (long long)B * N * H;
  const int blocks = (int)((total + 127) / 128 < 4096 ? (total + 127) / 128 : 4096);
  const size_t smem = (size_t)(W * W + H * W + W + H) * sizeof(float);
  model_head_kernel<<<blocks, 128, smem, st>>>(f0, f1, w0, b0, w2, b2, out, B, N, W, H);
  SG_LAUNCH_CHECK("model_head_kernel");
  return 0;
}

int launch_fold_in(const float* w_in, float* w_out, int d, int W, int chain, int kfirst, int nk,
                   cudaStream_t st) {
  fold_in_kernel<<<d, 64, 0, st>>>(w_in, w_out, d, W, chain, kfirst, nk);
  SG_LAUNCH_CHECK("fold_in_kernel");
  return 0;
}

// iffted[b,k,n,t] = sum_f ic[0][f][t] re[(b,n)][k*T+f] + ic[1][f][t] im[(b,n)][k*T+f]
// (stage-level API only; the fused path folds this into the output map).  act3: (R, 8T) = [re | im].
__global__ void irfft_rows_kernel(const float* __restrict__ act3, const float* __restrict__ ic,
                                  float* __restrict__ iffted, int B, int N, int T) {
  const long long total = (long long)B * 4 * N * T;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(idx % T), n = (int)((idx / T) % N), k = (int)((idx / ((long long)T * N)) % 4);
    const int b = (int)(idx / ((long long)T * N * 4));
    const float* re = act3 + ((long long)b * N + n) * (8 * T) + k * T;
    const float* im = re + 4 * T;
    float acc = 0.f;
    for (int f = 0; f <= T / 2; ++f) {
      acc = fmaf(ic[f * T + t], re[f], acc);
      acc = fmaf(ic[(T + f) * T + t], im[f], acc);
    }
    iffted[idx] = acc;
  }
}

int launch_irfft_rows(const float* act3, const float* ic, float* iffted, int B, int N, int T,
                      cudaStream_t st) {
  const long long total = (long long)B * 4 * N * T;
  const int blocks = (int)((total + 255) / 256 < 16384 ? (total + 255) / 256 : 16384);
  irfft_rows_kernel<<<blocks, 256, 0, st>>>(act3, ic, iffted, B, N, T);
  SG_LAUNCH_CHECK("irfft_rows_kernel");
  return 0;
}

// gfted (B,4,N,W) -> G4 (B*N, 4W) with column k*W+t   (stage-level API only)
__global__ void gfted_to_rows_kernel(const float* __restrict__ gfted, float* __restrict__ G4, int B,
                                     int N, int W) {
  const long long total = (long long)B * 4 * N * W;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(idx % W), n = (int)((idx / W) % N), k = (int)((idx / ((long long)W * N)) % 4);
    const int b = (int)(idx / ((long long)W * N * 4));
    G4[((long long)b * N + n) * (4 * W) + k * W + t] = gfted[idx];
  }
}

int launch_gfted_to_rows(const float* gfted, float* G4, int B, int N, int W, cudaStream_t st) {
  const long long total = (long long)B * 4 * N * W;
  const int blocks = (int)((total + 255) / 256 < 16384 ? (total + 255) / 256 : 16384);
  gfted_to_rows_kernel<<<blocks, 256, 0, st>>>(gfted, G4, B, N, W);
  SG_LAUNCH_CHECK("gfted_to_rows_kernel");
  return 0;
}

int launch_irfft_table(float* ic, int T, cudaStream_t st) {
  irfft_table_kernel<<<ceil_div(2 * T * T, 256), 256, 0, st>>>(ic, T);
  SG_LAUNCH_CHECK("irfft_table_kernel");
  return 0;
}

}  // namespace sg
