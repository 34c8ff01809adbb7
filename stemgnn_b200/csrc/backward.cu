// Backward of the hot path: what autograd does for the reference when handler.py:164 calls
// loss.backward() through base_model.py:136-179, restated as explicit kernels.
//
// Structure (reverse of api.cu's forward):  model head -> block 1 -> block 0 (each: head, folded
// output map, 3 GLU layers per chain, graph-Fourier contraction) -> Chebyshev stack -> Laplacian ->
// softmax attention -> key/query -> BPTT through the GRU -> input projection.
// Weight gradients are split-K fp32 GEMMs that ACCUMULATE into the caller's buffers (atomics), bias
// gradients are column sums; the folded DFT weights are differentiated in folded form and unfolded
// with the transposed twiddle tables.  The dropout mask is regenerated from (seed, offset).
#include "common.cuh"
#include "gemm.cuh"
#include "internal.cuh"

namespace sg {

// ---- generic helpers --------------------------------------------------------------------------------
template <bool AKM, bool BNK>
static int gemm(cudaStream_t st, int M, int N, int K, float alpha, const float* A, int lda, const float* B,
                int ldb, float beta, float* C, int ldc, const char* name) {
  GemmOperands g = {A, lda, 0, B, ldb, 0, nullptr, M, N, K, 0};
  EpiAxpby epi = {C, ldc, 0, beta != 0.f ? C : nullptr, ldc, 0, alpha, beta};
  return launch_sgemm<AKM, BNK, false>(g, epi, 1, st, name);
}
// C += A B   (split-K, atomic accumulation: the gradient-buffer contract of the ABI)
template <bool AKM, bool BNK>
static int gemm_acc(cudaStream_t st, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
                    float* C, int ldc, const char* name) {
  if (C == nullptr) return 0;
  GemmOperands g = {A, lda, 0, B, ldb, 0, nullptr, M, N, K, pick_ksplit(M, N, K)};
  EpiAtomicAdd epi = {C, ldc, 1.f};
  return launch_sgemm<AKM, BNK, false>(g, epi, 1, st, name);
}

// out[c] += scale * sum_r X[r*ld + c]
__global__ void __launch_bounds__(256) colsum_acc_kernel(const float* __restrict__ X, int rows, int cols,
                                                         int ld, float scale, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int r0 = blockIdx.y * 1024;
  const int r1 = min(rows, r0 + 1024);
  float acc = 0.f;
  if (c < cols)
    for (int r = r0 + (threadIdx.x >> 5); r < r1; r += 8) acc += X[(long long)r * ld + c];
  red[threadIdx.x >> 5][threadIdx.x & 31] = acc;
  __syncthreads();
  if (threadIdx.x < 32 && c < cols) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
    atomicAdd(out + c, scale * s);
  }
}
static int colsum_acc(cudaStream_t st, const float* X, int rows, int cols, int ld, float scale, float* out) {
  if (out == nullptr || rows <= 0 || cols <= 0) return 0;
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 1024));
  colsum_acc_kernel<<<grid, 256, 0, st>>>(X, rows, cols, ld, scale, out);
  SG_LAUNCH_CHECK("colsum_acc_kernel");
  return 0;
}

// ---- model head backward (base_model.py:174-179) --------------------------------------------------------
// forecast[b][h][n] = fc2(leaky(fc0(f0+f1)));  writes d_fsum (R,W), and row-major copies for the GEMMs
__global__ void __launch_bounds__(128) model_head_bwd_kernel(
    const float* __restrict__ f0, const float* __restrict__ f1, const float* __restrict__ w0,
    const float* __restrict__ b0, const float* __restrict__ w2, const float* __restrict__ d_out,
    float* __restrict__ fsum, float* __restrict__ act, float* __restrict__ d_hj, float* __restrict__ d_out_rows,
    float* __restrict__ d_fsum, int B, int N, int W, int H) {
  const long long R = (long long)B * N;
  for (long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x; row < R;
       row += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(row / N), n = (int)(row % N);
    float fs[64], dh[64];
    for (int t = 0; t < W; ++t) {
      fs[t] = f0[row * W + t] + f1[row * W + t];
      fsum[row * W + t] = fs[t];
    }
    float dout[64];
    for (int h = 0; h < H; ++h) {
      dout[h] = d_out[((long long)b * H + h) * N + n];
      d_out_rows[row * H + h] = dout[h];
    }
    for (int j = 0; j < W; ++j) {
      float hj = b0[j];
      for (int t = 0; t < W; ++t) hj = fmaf(fs[t], w0[j * W + t], hj);
      float da = 0.f;
      for (int h = 0; h < H; ++h) da = fmaf(dout[h], w2[h * W + j], da);
      act[row * W + j] = leaky_(hj, 0.01f);
      dh[j] = da * (hj >= 0.f ? 1.f : 0.01f);
      d_hj[row * W + j] = dh[j];
    }
    for (int t = 0; t < W; ++t) {
      float acc = 0.f;
      for (int j = 0; j < W; ++j) acc = fmaf(dh[j], w0[j * W + t], acc);
      d_fsum[row * W + t] = acc;
    }
  }
}

// ---- block head backward (base_model.py:68-72) ------------------------------------------------------------
struct HeadBwdArgs {
  const float* d_forecast;   // (R, W)
  const float* d_bc;         // (R, W) grad of the backcast output (block 0) or null
  const float* fs;           // (R, T) saved forecast_source
  const float* bc;           // (R, W) saved backcast (block 0)
  const float* wfr;          // (W, T)
  float* d_pre;              // (R, PW)
  float* negdz;              // (R, W) = -d(backcast pre-activation)  (block 0)
  int R, T, W, PW;
};
__global__ void __launch_bounds__(256) block_head_bwd_kernel(HeadBwdArgs a) {
  const long long total = (long long)a.R * a.PW;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int row = (int)(idx / a.PW), c = (int)(idx % a.PW);
    if (c < a.T) {
      float dfs = 0.f;
      for (int o = 0; o < a.W; ++o) dfs = fmaf(a.d_forecast[(long long)row * a.W + o], a.wfr[o * a.T + c], dfs);
      const float f = a.fs[(long long)row * a.T + c];
      a.d_pre[idx] = dfs * f * (1.f - f);
    } else {
      const int o = c - a.T;
      const float b = a.bc[(long long)row * a.W + o];
      const float dz = (a.d_bc != nullptr ? a.d_bc[(long long)row * a.W + o] : 0.f) * b * (1.f - b);
      a.d_pre[idx] = dz;
      a.negdz[(long long)row * a.W + o] = -dz;
    }
  }
}

// out[c*ldo + r] = in[r*ldi + c]   (32x32 shared-memory tiles)
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, int rows, int cols, int ldi,
                                                        float* __restrict__ out, int ldo) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int rr = ty; rr < 32; rr += 8)
    tile[rr][tx] = (r0 + rr < rows && c0 + tx < cols) ? in[(long long)(r0 + rr) * ldi + c0 + tx] : 0.f;
  __syncthreads();
  for (int cc = ty; cc < 32; cc += 8)
    if (c0 + cc < cols && r0 + tx < rows) out[(long long)(c0 + cc) * ldo + r0 + tx] = tile[tx][cc];
}
static int transpose(cudaStream_t st, const float* in, int rows, int cols, int ldi, float* out, int ldo) {
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32));
  transpose_kernel<<<grid, 256, 0, st>>>(in, rows, cols, ldi, out, ldo);
  SG_LAUNCH_CHECK("transpose_kernel");
  return 0;
}

// GLU gate backward fused with (i) the bias gradients (column sums of dl / dr) and (ii) the transposed
// copy dlrT (2N, ldT) that the tensor-core weight-gradient GEMM consumes as its K-major A operand.
__global__ void __launch_bounds__(256) glu_gate_bwd_fused_kernel(
    const float* __restrict__ d_out, int ldd, const float* __restrict__ l, const float* __restrict__ s, int R,
    int N, float* __restrict__ dlr, float* __restrict__ dlrT, int ldT, float* __restrict__ dbl,
    float* __restrict__ dbr) {
  __shared__ float tl[32][33], tr[32][33];
  __shared__ float cs[2][8][32];
  const int n0 = blockIdx.x * 32;
  const int rbase = blockIdx.y * 256;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = n0 + tx;
  float sl = 0.f, sr = 0.f;
  for (int sub = 0; sub < 8; ++sub) {
    const int r0 = rbase + sub * 32;
    if (r0 >= R) break;
    for (int rr = ty; rr < 32; rr += 8) {
      const int row = r0 + rr;
      float dl = 0.f, dr = 0.f;
      if (row < R && n < N) {
        const float dv = d_out[(long long)row * ldd + n];
        const float sv = s[(long long)row * N + n], lv = l[(long long)row * N + n];
        dl = dv * sv;
        dr = dv * lv * sv * (1.f - sv);
        dlr[(long long)row * 2 * N + n] = dl;
        dlr[(long long)row * 2 * N + N + n] = dr;
      }
      tl[rr][tx] = dl;
      tr[rr][tx] = dr;
      sl += dl;
      sr += dr;
    }
    __syncthreads();
    if (dlrT != nullptr) {
      for (int cc = ty; cc < 32; cc += 8) {
        if (n0 + cc < N && r0 + tx < R) {
          dlrT[(long long)(n0 + cc) * ldT + r0 + tx] = tl[tx][cc];
          dlrT[(long long)(N + n0 + cc) * ldT + r0 + tx] = tr[tx][cc];
        }
      }
    }
    __syncthreads();
  }
  cs[0][ty][tx] = sl;
  cs[1][ty][tx] = sr;
  __syncthreads();
  if (ty == 0 && n < N) {
    float a = 0.f, c = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { a += cs[0][i][tx]; c += cs[1][i][tx]; }
    if (dbl != nullptr) atomicAdd(dbl + n, a);
    if (dbr != nullptr) atomicAdd(dbr + n, c);
  }
}

// ---- unfold the gradient of DFT-folded first-layer weights (transpose of fold_in_kernel) -----------------------
// d_w[o][(kp+1)*W + f] += sum_t d_wf[o][kp*W + t] * tw(chain, f, t)
__global__ void fold_in_bwd_kernel(const float* __restrict__ d_wf, float* __restrict__ d_w, int d, int W,
                                   int chain) {
  const int o = blockIdx.x;
  for (int c = threadIdx.x; c < 3 * W; c += blockDim.x) {
    const int kp = c / W, f = c % W;
    float acc = 0.f;
    for (int t = 0; t < W; ++t) {
      const int ph = (f * t) % W;
      const float ang = 2.0f * (float)ph / (float)W;
      const float tw = chain == 0 ? cospif(ang) : -sinpif(ang);
      acc = fmaf(d_wf[(long long)o * 3 * W + kp * W + t], tw, acc);
    }
    d_w[(long long)o * 4 * W + (kp + 1) * W + f] += acc;
  }
}

// ---- graph-Fourier contraction backward --------------------------------------------------------------------------
// d_G (R, 3W) [(b,n)][(k',t)]  ->  dGp (3N, B*W) [(k',n)][(b,t)]
__global__ void permute_dg_kernel(const float* __restrict__ dG, float* __restrict__ dGp, int B, int N, int W) {
  const long long total = (long long)B * N * 3 * W;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    // idx enumerates the DESTINATION (coalesced writes): ((k'*N + n) * B + b) * W + t
    const int t = (int)(idx % W), b = (int)((idx / W) % B);
    const long long kn = idx / ((long long)W * B);
    const int n = (int)(kn % N), kp = (int)(kn / N);
    dGp[idx] = dG[((long long)b * N + n) * 3 * W + kp * W + t];
  }
}
// C(m, c=(b,t)) accumulated into d_x[(b*N + m) * W + t]
struct EpiScatterAccBNW {
  float* dx; int N, W; int accumulate;
  __device__ __forceinline__ void store4(int, int m, int n, int valid, float4 v, float4) const {
    const float vals[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < valid) {
        const int c = n + j, b = c / W, t = c - b * W;
        float* p = dx + ((long long)b * N + m) * W + t;
        *p = accumulate ? *p + vals[j] : vals[j];
      }
    }
  }
};

// ---- Laplacian backward (base_model.py:140-147) ---------------------------------------------------------------------
// one CTA per row i.  Inputs: dL (N,N), a_raw, deg, optional d_attention.  Output: dA_raw (N,N) pieces:
//   dAsym[i][j] = -dL[i][j] dh_i dh_j (+ d_attention[i][j]);   ddeg[i] (incl. diagonal and the dhat chain)
__global__ void __launch_bounds__(256) laplacian_bwd_rows_kernel(const float* __restrict__ dL,
                                                                 const float* __restrict__ a_raw,
                                                                 const float* __restrict__ deg,
                                                                 const float* __restrict__ d_att,
                                                                 float* __restrict__ dAsym,
                                                                 float* __restrict__ ddeg, int N) {
  __shared__ float red[32];
  const int i = blockIdx.x;
  const float degi = deg[i];
  const float sq = sqrtf(degi);
  const float dhi = 1.0f / (sq + 1e-7f);
  float acc = 0.f;   // d(dhat_i) = sum_j (dL[i][j] inner[i][j] + dL[j][i] inner[j][i]) dhat_j
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    const float dhj = 1.0f / (sqrtf(deg[j]) + 1e-7f);
    const float asym = 0.5f * (a_raw[(long long)i * N + j] + a_raw[(long long)j * N + i]);
    const float g_ij = dL[(long long)i * N + j], g_ji = dL[(long long)j * N + i];
    const float inner_ij = ((i == j) ? degi : 0.f) - asym;
    const float inner_ji = ((i == j) ? deg[j] : 0.f) - asym;
    acc += (g_ij * inner_ij + g_ji * inner_ji) * dhj;
    float da = -g_ij * dhi * dhj;
    if (d_att != nullptr) da += d_att[(long long)i * N + j];
    dAsym[(long long)i * N + j] = da;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) {
    const float g_ii = dL[(long long)i * N + i];
    float dd = g_ii * dhi * dhi;                                   // diagonal: inner_ii contains +deg_i
    dd += acc * (-dhi * dhi) * (0.5f / fmaxf(sq, 1e-30f));         // dhat = 1/(sqrt(deg)+eps)
    ddeg[i] = dd;
  }
}
// dA_raw[i][j] = 0.5 (dAsym[i][j] + dAsym[j][i]) + ddeg[i]
__global__ void laplacian_bwd_combine_kernel(const float* __restrict__ dAsym, const float* __restrict__ ddeg,
                                             float* __restrict__ dA, int N) {
  const long long total = (long long)N * N;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / N), j = (int)(idx % N);
    dA[idx] = 0.5f * (dAsym[idx] + dAsym[(long long)j * N + i]) + ddeg[i];
  }
}

// ---- attention backward (base_model.py:156-161) -----------------------------------------------------------------------
struct AttnBwdArgs {
  const float* key; const float* query; const float* row_m; const float* row_zinv;
  const float* dA;          // (N,N) grad of the batch-mean (dropped) attention
  float* dots;              // (B,N)  sum_j p * dp
  float* d_key;             // (B,N)
  float* d_query;           // (B,N)
  const uint8_t* mask; uint64_t seed, offset;
  const unsigned long long* offset_dev;
  float alpha, p; int use_dropout; int B, N;
};
__device__ __forceinline__ float attn_dp(const AttnBwdArgs& a, int b, int i, int j, float scale) {
  bool keep = true;
  if (a.use_dropout) {
    const uint64_t lin = ((uint64_t)b * a.N + i) * a.N + j;
    keep = a.mask != nullptr ? (a.mask[lin] != 0)
                             : dropout_keep(a.seed, a.offset + (a.offset_dev ? (uint64_t)__ldg(a.offset_dev) : 0ull), lin, a.p);
  }
  return keep ? a.dA[(long long)i * a.N + j] * scale : 0.f;
}
// pass 1: one warp per (b,i): dot = sum_j p dp;  d_key[b,i] = sum_j p (dp - dot) lrelu'
__global__ void __launch_bounds__(256) attention_bwd_rows_kernel(AttnBwdArgs a) {
  const int lane = threadIdx.x & 31;
  const int wrow = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (wrow >= a.B * a.N) return;
  const int b = wrow / a.N, i = wrow % a.N;
  const float scale = (a.use_dropout ? 1.0f / (1.0f - a.p) : 1.0f) / (float)a.B;
  const float ki = a.key[wrow], m = a.row_m[wrow], zi = a.row_zinv[wrow];
  float dot = 0.f;
  for (int j = lane; j < a.N; j += 32) {
    const float p = expf(leaky_(ki + a.query[(long long)b * a.N + j], a.alpha) - m) * zi;
    dot += p * attn_dp(a, b, i, j, scale);
  }
  dot = warp_sum(dot);
  float dk = 0.f;
  for (int j = lane; j < a.N; j += 32) {
    const float data = ki + a.query[(long long)b * a.N + j];
    const float p = expf(leaky_(data, a.alpha) - m) * zi;
    dk += p * (attn_dp(a, b, i, j, scale) - dot) * (data >= 0.f ? 1.f : a.alpha);
  }
  dk = warp_sum(dk);
  if (lane == 0) {
    a.dots[wrow] = dot;
    a.d_key[wrow] = dk;
  }
}
// pass 2: one warp per (b,j): d_query[b,j] = sum_i p(b,i,j) (dp - dot[b,i]) lrelu'
__global__ void __launch_bounds__(256) attention_bwd_cols_kernel(AttnBwdArgs a) {
  const int lane = threadIdx.x & 31;
  const int wcol = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (wcol >= a.B * a.N) return;
  const int b = wcol / a.N, j = wcol % a.N;
  const float scale = (a.use_dropout ? 1.0f / (1.0f - a.p) : 1.0f) / (float)a.B;
  const float qj = a.query[wcol];
  float dq = 0.f;
  for (int i = lane; i < a.N; i += 32) {
    const long long bi = (long long)b * a.N + i;
    const float data = a.key[bi] + qj;
    const float p = expf(leaky_(data, a.alpha) - a.row_m[bi]) * a.row_zinv[bi];
    dq += p * (attn_dp(a, b, i, j, scale) - a.dots[bi]) * (data >= 0.f ? 1.f : a.alpha);
  }
  dq = warp_sum(dq);
  if (lane == 0) a.d_query[wcol] = dq;
}

// ---- key/query contraction backward: d_wk[s] += sum_{b,i} h_s[b,i] d_key[b,i] ------------------------------------------------
__global__ void __launch_bounds__(256) keyquery_bwd_kernel(const float* __restrict__ h_all,
                                                           const float* __restrict__ d_key,
                                                           const float* __restrict__ d_query, int BN,
                                                           float* __restrict__ d_wk, float* __restrict__ d_wq) {
  __shared__ float red[32];
  const int s = blockIdx.x;
  const float* h = h_all + (long long)s * BN;
  float ak = 0.f, aq = 0.f;
  for (int i = threadIdx.x; i < BN; i += blockDim.x) {
    const float hv = h[i];
    ak = fmaf(hv, d_key[i], ak);
    aq = fmaf(hv, d_query[i], aq);
  }
  ak = block_sum(ak, red);
  aq = block_sum(aq, red);
  if (threadIdx.x == 0) {
    if (d_wk != nullptr) d_wk[s] += ak;
    if (d_wq != nullptr) d_wq[s] += aq;
  }
}

// ---- BPTT: one launch per step (reverse order) -----------------------------------------------------------------------------------
// grid = (ceil(N/32) k-tiles, ceil(B/4) batch groups), 256 threads.
//   dh      = dh_in[b][u] + d_key[b][u] wk[s] + d_query[b][u] wq[s]
//   h_s     = (1-z) n + z h_{s-1};  n = tanh(gi_n + r hn);  hn = W_hn h_{s-1} + b_hn
//   dh_out[b][k] = dh z  +  sum_rows d_gh[b][row] W_hh[row][k]
struct GruBwdArgs {
  const float* w_hh; const float* wk; const float* wq;
  const float* d_key; const float* d_query;
  const float* h_all; const float* g_r; const float* g_z; const float* g_n; const float* g_hn;
  float* dgh; float* dgi;            // (S*B, 3N)
  int B, N;
};
__global__ void __launch_bounds__(256) gru_bwd_step_kernel(GruBwdArgs a, int s, const float* __restrict__ dh_in,
                                                           float* __restrict__ dh_out) {
  extern __shared__ float sm[];
  const int N = a.N, B = a.B;
  float* sgh = sm;                   // [4][3N]
  float* sdz = sgh + 4 * 3 * N;      // [4][N]   dh * z
  float* part = sdz + 4 * N;         // [8][4][32]
  const int k0 = blockIdx.x * 32, b0 = blockIdx.y * 4;
  const float wk_s = a.wk[s], wq_s = a.wq[s];
  const long long step = (long long)s * B;
  for (int idx = threadIdx.x; idx < 4 * N; idx += blockDim.x) {
    const int bb = idx / N, u = idx - bb * N;
    const int b = b0 + bb;
    float d_r = 0.f, d_z = 0.f, d_np = 0.f, d_hn = 0.f, dz_dir = 0.f;
    if (b < B) {
      const long long e = (step + b) * N + u;
      const long long bu = (long long)b * N + u;
      const float dh = dh_in[bu] + a.d_key[bu] * wk_s + a.d_query[bu] * wq_s;
      const float r = a.g_r[e], z = a.g_z[e], n = a.g_n[e], hn = a.g_hn[e];
      const float hp = s > 0 ? a.h_all[e - (long long)B * N] : 0.f;
      const float dn = dh * (1.f - z);
      d_np = dn * (1.f - n * n);
      d_z = dh * (hp - n) * z * (1.f - z);
      d_r = d_np * hn * r * (1.f - r);
      d_hn = d_np * r;
      dz_dir = dh * z;
      if (blockIdx.x == 0) {
        float* gh = a.dgh + (step + b) * 3 * N;
        float* gi = a.dgi + (step + b) * 3 * N;
        gh[u] = d_r; gh[N + u] = d_z; gh[2 * N + u] = d_hn;
        gi[u] = d_r; gi[N + u] = d_z; gi[2 * N + u] = d_np;
      }
    }
    sgh[bb * 3 * N + u] = d_r;
    sgh[bb * 3 * N + N + u] = d_z;
    sgh[bb * 3 * N + 2 * N + u] = d_hn;
    sdz[bb * N + u] = dz_dir;
  }
  __syncthreads();
  const int kk = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int k = k0 + kk;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (k < N) {
    // 8 independent W_hh loads in flight per thread (the loop is L2-latency bound otherwise)
    for (int row0 = sl; row0 < 3 * N; row0 += 64) {
      float wv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int row = row0 + 8 * j;
        wv[j] = row < 3 * N ? __ldg(a.w_hh + (long long)row * N + k) : 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int row = min(row0 + 8 * j, 3 * N - 1);
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) acc[bb] = fmaf(sgh[bb * 3 * N + row], wv[j], acc[bb]);
      }
    }
  }
#pragma unroll
  for (int bb = 0; bb < 4; ++bb) part[(sl * 4 + bb) * 32 + kk] = acc[bb];
  __syncthreads();
  if (threadIdx.x < 128) {
    const int bb = threadIdx.x >> 5;
    const int b = b0 + bb;
    if (b < B && k < N) {
      float v = sdz[bb * N + k];
#pragma unroll
      for (int i = 0; i < 8; ++i) v += part[(i * 4 + bb) * 32 + kk];
      dh_out[(long long)b * N + k] = v;
    }
  }
}

// d_x[b][t][n] = d_xs[(n*B + b)*W + t] + d_x0[(b*N + n)*W + t]
__global__ void assemble_dx_kernel(const float* __restrict__ d_xs, const float* __restrict__ d_x0,
                                   float* __restrict__ d_x, int B, int W, int N) {
  const long long total = (long long)B * W * N;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(idx % N), t = (int)((idx / N) % W), b = (int)(idx / ((long long)N * W));
    d_x[idx] = d_xs[((long long)n * B + b) * W + t] + d_x0[((long long)b * N + n) * W + t];
  }
}

// y += a * x
__global__ void axpy_kernel(const float* __restrict__ x, float a, float* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] = fmaf(a, x[i], y[i]);
}

static inline int nblocks(long long total, int per, int cap) {
  long long b = (total + per - 1) / per;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

// ---- one spectral block ------------------------------------------------------------------------------------------------------------
static int block_backward(const stemgnn_dims_t& dm, const stemgnn_block_params_t& bp,
                          const stemgnn_block_grads_t& gr, int stack_idx, const float* x_bnw,
                          const float* x_bwn, const float* mul_L, const BlockWs& b, const BwdWs& w,
                          const float* d_forecast, const float* d_bc, float* d_xin, float* d_mul_L,
                          int first_mulL_writer, int gemm_mode, cudaStream_t st) {
  const int B = dm.B, N = dm.N, W = dm.W, T = dm.multi * W, d = 4 * T, R = B * N;
  const int PW = (stack_idx == 0) ? T + W : T;
  // (1) head
  HeadBwdArgs h = {d_forecast, d_bc, b.fs, b.bc_bnw, bp.forecast_result_w, w.d_pre, w.negdz, R, T, W, PW};
  block_head_bwd_kernel<<<nblocks((long long)R * PW, 256, 16384), 256, 0, st>>>(h);
  SG_LAUNCH_CHECK("block_head_bwd_kernel");
  SG_TRY((gemm_acc<true, false>(st, W, T, R, d_forecast, W, b.fs, T, gr.forecast_result_w, T, "d_forecast_result_w")));
  SG_TRY(colsum_acc(st, d_forecast, R, W, W, 1.f, gr.forecast_result_b));
  SG_TRY(colsum_acc(st, w.d_pre, R, T, PW, 1.f, gr.forecast_b));
  bool xin_written = false;
  if (stack_idx == 0) {
    SG_TRY(colsum_acc(st, w.d_pre + T, R, W, PW, 1.f, gr.backcast_b));
    SG_TRY((gemm_acc<true, false>(st, W, W, R, w.negdz, W, x_bnw, W, gr.shortcut_w, W, "d_shortcut_w")));
    SG_TRY(colsum_acc(st, w.negdz, R, W, W, 1.f, gr.shortcut_b));
    if (d_xin != nullptr) {   // d_x += (-dz) @ Wsc
      SG_TRY((gemm<false, false>(st, R, W, W, 1.f, w.negdz, W, bp.shortcut_w, W, 0.f, d_xin, W, "d_x_shortcut")));
      xin_written = true;
    }
  }
  // (2) folded output map: pre = act3 @ woutT^T,  woutT (PW, 8T) = [Wf; Wb] @ RI^T
  SG_TRY((gemm<false, false>(st, R, 2 * d, PW, 1.f, w.d_pre, PW, b.wout, 2 * d, 0.f, w.d_act3, 2 * d, "d_act3")));
  SG_CUDA(cudaMemsetAsync(w.d_wout, 0, (size_t)8 * T * PW * sizeof(float), st));
  SG_TRY((gemm_acc<true, false>(st, PW, 2 * d, R, w.d_pre, PW, b.act3, 2 * d, w.d_wout, 2 * d, "d_woutT")));
  SG_TRY((gemm_acc<false, false>(st, T, T, 8 * T, w.d_wout, 8 * T, b.ri, T, gr.forecast_w, T, "d_forecast_w")));
  SG_TRY((gemm<true, false>(st, 8 * T, T, T, 1.f, w.d_wout, 8 * T, bp.forecast_w, T, 0.f, w.d_ri, T, "d_ri_f")));
  if (stack_idx == 0) {
    SG_TRY((gemm_acc<false, false>(st, W, T, 8 * T, w.d_wout + (size_t)T * 8 * T, 8 * T, b.ri, T, gr.backcast_w, T,
                                  "d_backcast_w")));
    SG_TRY((gemm<true, false>(st, 8 * T, T, W, 1.f, w.d_wout + (size_t)T * 8 * T, 8 * T, bp.backcast_w, T, 1.f,
                              w.d_ri, T, "d_ri_b")));
  }
  //   RI[c][k*T+f][u] = sum_t ic[c][f][t] weight[k][t][u]  ->  d_weight[k][t][u] += sum_{c,f} ic[c][f][t] d_RI
  if (gr.weight != nullptr) {
    for (int c = 0; c < 2; ++c) {
      GemmOperands g = {b.ic + (size_t)c * T * T, T, 0, w.d_ri + (size_t)c * 4 * T * T, T, (long long)T * T,
                        nullptr, T, T, T, 0};
      EpiAxpby epi = {gr.weight, T, (long long)T * T, gr.weight, T, (long long)T * T, 1.f, 1.f};
      SG_TRY((launch_sgemm<true, false, false>(g, epi, 4, st, "d_weight")));
    }
  }
  // (3) GLU layers 3 -> 1, both chains;  d_G accumulates over the chains
  const int ncol = 3 * W;
  SG_CUDA(cudaMemsetAsync(w.d_w1f, 0, (size_t)4 * d * ncol * sizeof(float), st));
  for (int c = 0; c < 2; ++c) {
    const float* d_out = w.d_act3 + (size_t)c * d;
    int ldd = 2 * d;
    for (int layer = 2; layer >= 0; --layer) {
      const int gidx = 2 * layer + c;
      const int ldT = (R + 3) / 4 * 4;
      const bool tc = (gemm_mode != 1) && layer > 0 && gr.glu_left_w[gidx] != nullptr &&
                      gr.glu_right_w[gidx] != nullptr;
      {
        dim3 ggrid(ceil_div(d, 32), ceil_div(R, 256));
        glu_gate_bwd_fused_kernel<<<ggrid, 256, 0, st>>>(d_out, ldd, b.save_l[gidx], b.save_s[gidx], R, d, w.dlr,
                                                        tc ? w.dlrT : nullptr, ldT, gr.glu_left_b[gidx],
                                                        gr.glu_right_b[gidx]);
        SG_LAUNCH_CHECK("glu_gate_bwd_fused_kernel");
      }
      if (layer > 0) {
        const float* in = (layer == 2 ? b.act2 : b.act1) + (size_t)c * R * d;
        float* d_in = w.d_act[layer & 1];
        int rc = -1;
        if (tc) {
          // weight gradients: [dWl; dWr] (2d x d) += dlrT (2d x R) . inT (d x R)^T   (K = R, split-K atomics)
          SG_TRY(transpose(st, in, R, d, d, w.inT, ldT));
          rc = tc_gemm(2 * d, d, R, 1.f, w.dlrT, ldT, w.inT, ldT, d, gr.glu_left_w[gidx], gr.glu_right_w[gidx], d, d,
                       d, 1, 36, st, gemm_mode == 0 ? -1 : 0);   // 4 row tiles x 36 K splits = 144 CTAs (one wave)
          if (rc > 0) return rc;
          if (rc == 0) {
            // input gradient: d_in (R x d) = dlr (R x 2d) . [Wl^T | Wr^T] (d x 2d)^T
            SG_TRY(transpose(st, bp.glu_left_w[gidx], d, d, d, w.wsT, 2 * d));
            SG_TRY(transpose(st, bp.glu_right_w[gidx], d, d, d, w.wsT + d, 2 * d));
            rc = tc_gemm(R, d, 2 * d, 1.f, w.dlr, 2 * d, w.wsT, 2 * d, d, d_in, nullptr, 0, d, d, 0, 1, st,
                         gemm_mode == 0 ? -1 : 0);
            if (rc > 0) return rc;
            SG_CHECK(rc == 0, "tcgen05 backward GEMM rejected after the weight-gradient GEMM ran");
          }
        }
        if (rc < 0) {
          SG_TRY((gemm_acc<true, false>(st, d, d, R, w.dlr, 2 * d, in, d, gr.glu_left_w[gidx], d, "d_glu_left_w")));
          SG_TRY((gemm_acc<true, false>(st, d, d, R, w.dlr + d, 2 * d, in, d, gr.glu_right_w[gidx], d, "d_glu_right_w")));
          SG_TRY((gemm<false, false>(st, R, d, d, 1.f, w.dlr, 2 * d, bp.glu_left_w[gidx], d, 0.f, d_in, d, "d_act_l")));
          SG_TRY((gemm<false, false>(st, R, d, d, 1.f, w.dlr + d, 2 * d, bp.glu_right_w[gidx], d, 1.f, d_in, d, "d_act_r")));
        }
        d_out = d_in;
        ldd = d;
      } else {
        float* dwl = w.d_w1f + (size_t)(c * 2 + 0) * d * ncol;
        float* dwr = w.d_w1f + (size_t)(c * 2 + 1) * d * ncol;
        SG_TRY((gemm_acc<true, false>(st, d, ncol, R, w.dlr, 2 * d, b.G, ncol, dwl, ncol, "d_w1f_l")));
        SG_TRY((gemm_acc<true, false>(st, d, ncol, R, w.dlr + d, 2 * d, b.G, ncol, dwr, ncol, "d_w1f_r")));
        const float* w1l = b.w1f + (size_t)(c * 2 + 0) * d * ncol;
        const float* w1r = b.w1f + (size_t)(c * 2 + 1) * d * ncol;
        SG_TRY((gemm<false, false>(st, R, ncol, d, 1.f, w.dlr, 2 * d, w1l, ncol, c == 0 ? 0.f : 1.f, w.d_G, ncol, "d_G_l")));
        SG_TRY((gemm<false, false>(st, R, ncol, d, 1.f, w.dlr + d, 2 * d, w1r, ncol, 1.f, w.d_G, ncol, "d_G_r")));
        if (gr.glu_left_w[c] != nullptr) {
          fold_in_bwd_kernel<<<d, 64, 0, st>>>(dwl, gr.glu_left_w[c], d, W, c);
          SG_LAUNCH_CHECK("fold_in_bwd_kernel");
        }
        if (gr.glu_right_w[c] != nullptr) {
          fold_in_bwd_kernel<<<d, 64, 0, st>>>(dwr, gr.glu_right_w[c], d, W, c);
          SG_LAUNCH_CHECK("fold_in_bwd_kernel");
        }
      }
    }
  }
  // (4) graph-Fourier contraction  G[(b,n)][(k',t)] = sum_m L_{k'+1}[n][m] x[b][m][t]
  permute_dg_kernel<<<nblocks((long long)R * ncol, 256, 16384), 256, 0, st>>>(w.d_G, w.d_Gp, B, N, W);
  SG_LAUNCH_CHECK("permute_dg_kernel");
  const size_t nn = (size_t)N * N;
  //   d_mul_L[k'+1] (+)= dGp (3N x BW) @ x_bwn (BW x N)
  SG_TRY((gemm<false, false>(st, 3 * N, N, B * W, 1.f, w.d_Gp, B * W, x_bwn, N, first_mulL_writer ? 0.f : 1.f,
                             d_mul_L + nn, N, "d_mul_L")));
  if (d_xin != nullptr) {   // d_x[(b,m)][t] (+)= sum_(k',n) L[(k',n)][m] dGp[(k',n)][(b,t)]
    GemmOperands g = {mul_L + nn, N, 0, w.d_Gp, B * W, 0, nullptr, N, B * W, 3 * N, 0};
    EpiScatterAccBNW epi = {d_xin, N, W, xin_written ? 1 : 0};
    SG_TRY((launch_sgemm<true, false, false>(g, epi, 1, st, "d_x_gft")));
  }
  return 0;
}

// ---- whole model --------------------------------------------------------------------------------------------------------------------
int model_backward(const stemgnn_dims_t* dims, const stemgnn_params_t* p, const stemgnn_fwd_opts_t* opts,
                   const float* x, const float* d_forecast, const float* d_attention,
                   const stemgnn_grads_t* grads, float* d_x, void* workspace, size_t workspace_bytes,
                   cudaStream_t st) {
  SG_CHECK(dims && p && opts && x && d_forecast && grads && workspace, "backward: null argument");
  SG_CHECK(opts->training, "backward needs the workspace of a training forward (opts.training = 1)");
  const stemgnn_dims_t& dm = *dims;
  SG_CHECK(workspace_bytes >= stemgnn_workspace_bytes(dims, 1), "backward: workspace too small");
  Workspace ws = carve_workspace(dm, 1, static_cast<float*>(workspace));
  const BwdWs& w = ws.bwd;
  const stemgnn_grads_t& gr = *grads;
  const int B = dm.B, N = dm.N, W = dm.W, H = dm.H, R = B * N;
  const size_t nn = (size_t)N * N;

  // ---- model head ----
  model_head_bwd_kernel<<<nblocks(R, 128, 4096), 128, 0, st>>>(
      ws.blk[0].forecast, ws.blk[1].forecast, p->fc0_w, p->fc0_b, p->fc2_w, d_forecast, w.h_fsum, w.h_act,
      w.h_dhj, w.h_dout, w.d_fsum, B, N, W, H);
  SG_LAUNCH_CHECK("model_head_bwd_kernel");
  SG_TRY((gemm_acc<true, false>(st, H, W, R, w.h_dout, H, w.h_act, W, gr.fc2_w, W, "d_fc2_w")));
  SG_TRY(colsum_acc(st, w.h_dout, R, H, H, 1.f, gr.fc2_b));
  SG_TRY((gemm_acc<true, false>(st, W, W, R, w.h_dhj, W, w.h_fsum, W, gr.fc0_w, W, "d_fc0_w")));
  SG_TRY(colsum_acc(st, w.h_dhj, R, W, W, 1.f, gr.fc0_b));

  // ---- blocks (reverse) ----
  SG_TRY(block_backward(dm, p->block[1], gr.block[1], 1, ws.blk[0].bc_bnw, ws.blk[0].bc_bwn, ws.mul_L,
                        ws.blk[1], w, w.d_fsum, nullptr, w.d_bc, w.d_mul_L, 1, opts->gemm_mode, st));
  SG_TRY(block_backward(dm, p->block[0], gr.block[0], 0, ws.x_bnw, x, ws.mul_L, ws.blk[0], w, w.d_fsum,
                        w.d_bc, d_x != nullptr ? w.d_x0 : nullptr, w.d_mul_L, 0, opts->gemm_mode, st));

  // ---- Chebyshev stack: L2 = 2 L L, L3 = 2 L L2 - L  (base_model.py:130-132) ----
  const float* L = ws.mul_L + nn;
  const float* L2 = ws.mul_L + 2 * nn;
  float* dL1 = w.d_mul_L + nn;
  float* dL2 = w.d_mul_L + 2 * nn;
  float* dL3 = w.d_mul_L + 3 * nn;
  SG_TRY((gemm<true, false>(st, N, N, N, 2.f, L, N, dL3, N, 1.f, dL2, N, "cheb_bwd_dL2")));   // dL2 += 2 L^T dL3
  axpy_kernel<<<nblocks((long long)nn, 256, 2048), 256, 0, st>>>(dL3, -1.f, dL1, (long long)nn);   // dL -= dL3
  SG_LAUNCH_CHECK("axpy_kernel");
  SG_TRY((gemm<false, true>(st, N, N, N, 2.f, dL3, N, L2, N, 1.f, dL1, N, "cheb_bwd_a")));     // dL += 2 dL3 L2^T
  SG_TRY((gemm<false, true>(st, N, N, N, 2.f, dL2, N, L, N, 1.f, dL1, N, "cheb_bwd_b")));      // dL += 2 dL2 L^T
  SG_TRY((gemm<true, false>(st, N, N, N, 2.f, L, N, dL2, N, 1.f, dL1, N, "cheb_bwd_c")));      // dL += 2 L^T dL2

  // ---- Laplacian (base_model.py:140-147) ----
  laplacian_bwd_rows_kernel<<<N, 256, 0, st>>>(dL1, ws.a_raw, ws.deg, d_attention, w.dAsym, w.ddeg, N);
  SG_LAUNCH_CHECK("laplacian_bwd_rows_kernel");
  laplacian_bwd_combine_kernel<<<nblocks((long long)nn, 256, 2048), 256, 0, st>>>(w.dAsym, w.ddeg, w.dA, N);
  SG_LAUNCH_CHECK("laplacian_bwd_combine_kernel");
  // global-batch graph (forward all-reduced the shard-mean attention): the loss of EVERY rank depends on this rank's
  // attention, so its gradient is the mean over ranks of the local d loss_k / d attention
  if (opts->graph_allreduce != nullptr) opts->graph_allreduce(w.dA, (long long)nn, opts->graph_allreduce_user, st);

  // ---- softmax attention (base_model.py:156-161) ----
  AttnBwdArgs ab = {};
  ab.key = ws.key; ab.query = ws.query; ab.row_m = ws.row_m; ab.row_zinv = ws.row_zinv; ab.dA = w.dA;
  ab.dots = w.dots; ab.d_key = w.d_key; ab.d_query = w.d_query;
  ab.mask = opts->dropout_mask; ab.seed = opts->dropout_seed; ab.offset = opts->dropout_offset;
  ab.offset_dev = opts->dropout_offset_dev;
  ab.alpha = opts->leaky_alpha; ab.p = opts->dropout_p;
  ab.use_dropout = (opts->training && (opts->dropout_p > 0.f || opts->dropout_mask != nullptr)) ? 1 : 0;
  ab.B = B; ab.N = N;
  attention_bwd_rows_kernel<<<ceil_div(R, 8), 256, 0, st>>>(ab);
  SG_LAUNCH_CHECK("attention_bwd_rows_kernel");
  attention_bwd_cols_kernel<<<ceil_div(R, 8), 256, 0, st>>>(ab);
  SG_LAUNCH_CHECK("attention_bwd_cols_kernel");

  // ---- key/query contraction (base_model.py:154-155) ----
  if (gr.weight_key != nullptr || gr.weight_query != nullptr) {
    keyquery_bwd_kernel<<<N, 256, 0, st>>>(ws.h_all, w.d_key, w.d_query, R, gr.weight_key, gr.weight_query);
    SG_LAUNCH_CHECK("keyquery_bwd_kernel");
  }

  // ---- BPTT through the GRU (base_model.py:137) ----
  GruBwdArgs ga = {p->gru_w_hh, p->weight_key, p->weight_query, w.d_key, w.d_query, ws.h_all,
                   ws.g_r, ws.g_z, ws.g_n, ws.g_hn, w.dgh, ws.gi /* reused as dgi */, B, N};
  int brc = gru_bwd_cluster(p->gru_w_hh, p->weight_key, p->weight_query, w.d_key, w.d_query, ws.h_all, ws.g_r,
                            ws.g_z, ws.g_n, ws.g_hn, w.dgh, ws.gi, B, N, st);
  if (brc > 0) return brc;
  if (brc < 0) {   // generic fallback: one launch per step, W_hh streamed from L2
    SG_CUDA(cudaMemsetAsync(w.dh[0], 0, (size_t)R * sizeof(float), st));
    const size_t smem = (size_t)(4 * 3 * N + 4 * N + 8 * 4 * 32) * sizeof(float);
    SG_CHECK(smem <= 200 * 1024, "gru backward: N=%d too large for the step kernel", N);
    static size_t smem_set_dev[64] = {};   // function attributes are per device
  int dev_ = 0;
  (void)cudaGetDevice(&dev_);
  size_t& smem_set = smem_set_dev[dev_ & 63];
    if (smem > smem_set) {
      SG_CUDA(cudaFuncSetAttribute(gru_bwd_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      smem_set = smem;
    }
    dim3 bgrid(ceil_div(N, 32), ceil_div(B, 4));
    int cur = 0;
    for (int s = N - 1; s >= 0; --s) {
      gru_bwd_step_kernel<<<bgrid, 256, smem, st>>>(ga, s, w.dh[cur], w.dh[cur ^ 1]);
      count_launch();
      cur ^= 1;
    }
    SG_LAUNCH_CHECK("gru_bwd_step_kernel");
  }
  const float* dgh = w.dgh;
  const float* dgi = ws.gi;
  const int SB = N * B;
  //   dW_hh += DGH[1:]^T h_all[:-1] ;  db_hh += colsum(DGH)
  {
    const int Kr = SB - B;                      // step 0 has h_{-1} = 0
    int rc = -1;
    if (opts->gemm_mode != 1 && gr.gru_w_hh != nullptr && Kr > 0) {
      // tcgen05 TF32: both operands made K-major by one transpose each, output columns in chunks of <= 256
      const int ldT = (Kr + 3) / 4 * 4;
      SG_TRY(transpose(st, dgh + (size_t)B * 3 * N, Kr, 3 * N, 3 * N, w.dghT, ldT));
      SG_TRY(transpose(st, ws.h_all, Kr, N, N, w.hT, ldT));
      rc = 0;
      for (int n0 = 0; n0 < N && rc == 0; n0 += 256) {
        const int nn_ = N - n0 < 256 ? N - n0 : 256;
        const int npad = (nn_ + 15) / 16 * 16;
        rc = tc_gemm(3 * N, npad, Kr, 1.f, w.dghT, ldT, w.hT + (size_t)n0 * ldT, ldT, nn_, gr.gru_w_hh + n0, nullptr,
                     0, N, nn_, 1, 16, st, opts->gemm_mode == 0 ? -1 : 0);
        if (rc > 0) return rc;
        // a chunk that is unsupported AFTER earlier chunks were accumulated must not fall back to the full fp32 product
        // (it would double-count the finished columns)
        SG_CHECK(!(rc < 0 && n0 > 0), "d_gru_w_hh: tensor-core column chunk at %d unsupported after earlier chunks ran", n0);
      }
      SG_CHECK(rc == 0 || rc == -1, "tc dW_hh");
    }
    if (rc < 0)
      SG_TRY((gemm_acc<true, false>(st, 3 * N, N, Kr, dgh + (size_t)B * 3 * N, 3 * N, ws.h_all, N, gr.gru_w_hh, N,
                                    "d_gru_w_hh")));
  }
  SG_TRY(colsum_acc(st, dgh, SB, 3 * N, 3 * N, 1.f, gr.gru_b_hh));
  //   dW_ih += DGI^T xs ;  db_ih += colsum(DGI) ;  d_xs = DGI W_ih
  SG_TRY((gemm_acc<true, false>(st, 3 * N, W, SB, dgi, 3 * N, ws.xs, W, gr.gru_w_ih, W, "d_gru_w_ih")));
  SG_TRY(colsum_acc(st, dgi, SB, 3 * N, 3 * N, 1.f, gr.gru_b_ih));
  if (d_x != nullptr) {
    SG_TRY((gemm<false, false>(st, SB, W, 3 * N, 1.f, dgi, 3 * N, p->gru_w_ih, W, 0.f, w.d_xs, W, "d_xs")));
    assemble_dx_kernel<<<nblocks((long long)R * W, 256, 8192), 256, 0, st>>>(w.d_xs, w.d_x0, d_x, B, W, N);
    SG_LAUNCH_CHECK("assemble_dx_kernel");
  }
  return 0;
}

}  // namespace sg
