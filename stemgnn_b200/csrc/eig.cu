// Fused Laplacian build + symmetric eigendecomposition (north_star: "one fused kernel for Laplacian build + symeig on
// N <= 512 nodes staged entirely in shared memory with warp-level Jacobi sweeps").
//
// Reference hooks: models/base_model.py:106-119 `get_laplacian`, :164-165 `graph_fft` (dead code in the reference — its
// forward uses the Chebyshev stack of :121-134 instead).  This kernel is therefore an OPT-IN path
// (`Model.graph_mode = "eig"`): L = U diag(lambda) U^T, from which the polynomial stack [0, L, 2L^2, 4L^3 - L] is
// rebuilt as U p(Lambda) U^T.  The default forward keeps the reference-faithful polynomial GEMMs.
//
// Algorithm: one-sided (Hestenes) Jacobi on the columns of G = L with accumulated rotations V (G = L V, so at
// convergence column j of G is lambda_j v_j).  One thread-block CLUSTER owns the whole problem:
//   * a warp owns one "seat" = a pair of columns (p, q) of G and V, held in REGISTERS (rows lane, lane+32, ...);
//   * per round every warp orthogonalises its pair: alpha = |g_p|^2, beta = |g_q|^2, gamma = g_p.g_q by warp shuffles,
//     one plane rotation applied to the four register columns;
//   * round-robin (circle) ordering: after a round every column moves to the neighbouring seat — through a shared-memory
//     mailbox of the destination warp, i.e. through DISTRIBUTED shared memory when the neighbour lives in the next CTA of
//     the cluster; n-1 rounds = one sweep, sweeps until max |gamma| / sqrt(alpha beta) < tol.
//   * the Laplacian D^(diag(deg) - (A + A^T)/2) D^ (base_model.py:141-147) is formed on the fly while the columns are
//     loaded, so L itself never round-trips through HBM.
// G and V never leave registers / shared memory between the first load and the final store.
#include <cooperative_groups.h>

#include "common.cuh"
#include "gemm.cuh"
#include "internal.cuh"

namespace cg = cooperative_groups;

namespace sg {
namespace {

struct EigArgs {
  const float* a_raw;    // (N,N) batch-mean attention, not symmetrised
  const float* deg;      // (N)   row sums of a_raw
  float* lambda;         // (n)   eigenvalue of output column j (unsorted; n = N rounded up to even)
  float* U;              // (n,n) row-major, column j = eigenvector j (row n-1 / one unit column are padding when N is odd)
  int* info;             // [0] sweeps used, [1] 1 if converged
  int N, n;              // real size, padded (even) size
  int max_sweeps;
  float tol;
};

// R = rows per lane (n <= 32 R), CS = cluster size
template <int R, int CS>
__global__ void __launch_bounds__(1024 / (R > 8 ? 2 : 1), 1) laplacian_eig_kernel(EigArgs a) {
  extern __shared__ __align__(16) float smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n = a.n, N = a.N, m = n >> 1;
  const int wpc = (m + CS - 1) / CS;                 // seats (warps) per CTA
  const int seat = rank * wpc + w;
  const bool active = w < wpc && seat < m;
  // mailbox of local seat w: [top G | top V | bot G | bot V], n floats each
  float* inbox = smem;
  float* conv = smem + (size_t)wpc * 4 * n;          // [0]: max rotation measure of the sweep (CTA 0's copy is authoritative)
  if (threadIdx.x == 0) conv[0] = 0.f;

  // ---- load: columns p = seat (top row), q = n-1-seat (bottom row) of L; V = I ------------------------------------
  float gp[R], gq[R], vp[R], vq[R];
  int p = seat, q = n - 1 - seat;
  auto lap = [&](int i, int j) -> float {
    if (i >= N || j >= N) return 0.f;
    const float asym = 0.5f * (__ldg(a.a_raw + (size_t)i * N + j) + __ldg(a.a_raw + (size_t)j * N + i));
    const float di = 1.0f / (sqrtf(__ldg(a.deg + i)) + 1e-7f);
    const float dj = 1.0f / (sqrtf(__ldg(a.deg + j)) + 1e-7f);
    const float inner = ((i == j) ? __ldg(a.deg + i) : 0.f) - asym;
    return di * (inner * dj);
  };
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int i = lane + 32 * r;
    gp[r] = gq[r] = vp[r] = vq[r] = 0.f;
    if (active && i < n) {
      gp[r] = lap(i, p);
      gq[r] = lap(i, q);
      vp[r] = i == p ? 1.f : 0.f;
      vq[r] = i == q ? 1.f : 0.f;
    }
  }
  cluster.sync();

  // destinations of the circle method (seats k = 0..m-1): top[0] stays; top[k] -> top[k+1]; top[m-1] -> bot[m-1];
  // bot[k] -> bot[k-1]; bot[0] -> top[1]
  auto slot_ptr = [&](int dseat, int which) -> float* {     // which: 0 = top, 1 = bottom half of the seat's mailbox
    const int dr = dseat / wpc, dw = dseat - dr * wpc;
    float* local = inbox + ((size_t)dw * 4 + 2 * which) * n;
    return dr == rank ? local : cluster.map_shared_rank(local, dr);
  };
  float* top_dst = nullptr;
  float* bot_dst = nullptr;
  if (active) {
    if (seat == 0) top_dst = nullptr;                            // fixed player
    else if (seat == m - 1) top_dst = slot_ptr(m - 1, 1);
    else top_dst = slot_ptr(seat + 1, 0);
    if (m == 1) bot_dst = nullptr;
    else if (seat == 0) bot_dst = slot_ptr(1, 0);
    else bot_dst = slot_ptr(seat - 1, 1);
  }
  const float* my_box = inbox + (size_t)w * 4 * n;

  int sweeps = 0, converged = 0;
  for (; sweeps < a.max_sweeps && !converged; ++sweeps) {
    float worst = 0.f;
    for (int round = 0; round < n - 1; ++round) {
      if (active) {
        float al = 0.f, be = 0.f, ga = 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          al = fmaf(gp[r], gp[r], al);
          be = fmaf(gq[r], gq[r], be);
          ga = fmaf(gp[r], gq[r], ga);
        }
        al = warp_sum(al); be = warp_sum(be); ga = warp_sum(ga);
        const float denom = sqrtf(al * be);
        // a column whose norm is at the fp32 noise level of its partner (|lambda_q| < 1e-5 |lambda_p|) carries no resolvable
        // direction any more: it counts as converged (L always has one eigenvalue ~ 0)
        const bool noise = fminf(al, be) <= 1e-10f * fmaxf(al, be);
        const float rel = (denom > 0.f && !noise) ? fabsf(ga) / denom : 0.f;
        worst = fmaxf(worst, rel);
        if (rel > 1e-9f && denom > 1e-30f) {
          const float zeta = (be - al) / (2.f * ga);
          const float t = copysignf(1.f, zeta) / (fabsf(zeta) + sqrtf(1.f + zeta * zeta));
          const float c = 1.0f / sqrtf(1.f + t * t), s = c * t;   // IEEE sqrt / div: rsqrtf's biased error makes the columns' norms drift
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const float g0 = gp[r], g1 = gq[r], v0 = vp[r], v1 = vq[r];
            gp[r] = c * g0 - s * g1; gq[r] = s * g0 + c * g1;
            vp[r] = c * v0 - s * v1; vq[r] = s * v0 + c * v1;
          }
        }
        // move the columns to their next seats (mailboxes; DSMEM across CTAs)
        if (m > 1) {
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const int i = lane + 32 * r;
            if (i < n) {
              if (top_dst != nullptr) { top_dst[i] = gp[r]; top_dst[n + i] = vp[r]; }
              bot_dst[i] = gq[r]; bot_dst[n + i] = vq[r];
            }
          }
        }
      }
      cluster.sync();                                   // every column has landed in its next seat
      if (active && m > 1) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int i = lane + 32 * r;
          if (i < n) {
            if (seat != 0) { gp[r] = my_box[i]; vp[r] = my_box[n + i]; }
            gq[r] = my_box[2 * n + i]; vq[r] = my_box[3 * n + i];
          }
        }
      }
      cluster.sync();                                   // mailboxes may be overwritten again
    }
    // convergence of the sweep: max over all seats, collected in CTA 0's shared memory
    if (active && lane == 0) {
      float* c0 = cluster.map_shared_rank(conv, 0);
      atomicMax(reinterpret_cast<int*>(c0), __float_as_int(worst));      // non-negative floats order like ints
    }
    cluster.sync();
    const float wmax = *cluster.map_shared_rank(conv, 0);
    converged = wmax < a.tol ? 1 : 0;
    cluster.sync();
    if (rank == 0 && threadIdx.x == 0) conv[0] = 0.f;
    cluster.sync();
  }

  // ---- store: lambda_j = v_j . g_j (Rayleigh quotient keeps the sign), U[:, j] = v_j ---------------------------------
  if (active) {
    float lp = 0.f, lq = 0.f, np2 = 0.f, nq2 = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      lp = fmaf(vp[r], gp[r], lp);
      lq = fmaf(vq[r], gq[r], lq);
      np2 = fmaf(vp[r], vp[r], np2);
      nq2 = fmaf(vq[r], vq[r], nq2);
    }
    lp = warp_sum(lp); lq = warp_sum(lq); np2 = warp_sum(np2); nq2 = warp_sum(nq2);
    lp /= np2; lq /= nq2;                              // Rayleigh quotient of the (re-normalised) accumulated rotation
    const float ip = 1.0f / sqrtf(np2), iq = 1.0f / sqrtf(nq2);
#pragma unroll
    for (int r = 0; r < R; ++r) { vp[r] *= ip; vq[r] *= iq; }
    const int jp = 2 * seat, jq = 2 * seat + 1;        // output slots (order is irrelevant: the host side sorts)
    if (lane == 0) { a.lambda[jp] = lp; a.lambda[jq] = lq; }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int i = lane + 32 * r;
      if (i < n) {
        a.U[(size_t)i * n + jp] = vp[r];
        a.U[(size_t)i * n + jq] = vq[r];
      }
    }
  }
  if (rank == 0 && threadIdx.x == 0) { a.info[0] = sweeps; a.info[1] = converged; }
  cluster.sync();
}

template <int R, int CS>
int launch_eig(const EigArgs& a, cudaStream_t st) {
  auto kern = laplacian_eig_kernel<R, CS>;
  const int m = a.n / 2;
  const int wpc = ceil_div(m, CS);
  const size_t smem = ((size_t)wpc * 4 * a.n + 4) * sizeof(float);
  if (smem > 227 * 1024 || wpc * 32 > 1024 / (R > 8 ? 2 : 1)) return -1;
  SG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (CS > 8) SG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CS);
  cfg.blockDim = dim3(wpc * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SG_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
  count_launch();
  return 0;
}

// out[k] = U diag(p_k(lambda)) : scaled copy used by the U p(Lambda) U^T reconstruction
__global__ void scale_columns_kernel(const float* __restrict__ U, const float* __restrict__ lambda, int n, int N, int term,
                                     float* __restrict__ out) {
  const long long total = (long long)N * n;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / n), j = (int)(idx % n);
    const float l = lambda[j];
    const float pk = term == 1 ? l : term == 2 ? 2.f * l * l : 4.f * l * l * l - l;
    out[idx] = U[(size_t)i * n + j] * pk;
  }
}

}  // namespace

int laplacian_eig(const float* a_raw, const float* deg, int N, float* lambda, float* U, int* info, int max_sweeps,
                  float tol, cudaStream_t st) {
  SG_CHECK(N >= 2 && N <= 512, "laplacian_eig: N=%d outside [2,512]", N);
  EigArgs a = {a_raw, deg, lambda, U, info, N, (N + 1) & ~1, max_sweeps, tol};
  int rc = -1;
  const int n = a.n;
  if (n <= 128) rc = launch_eig<4, 8>(a, st);
  else if (n <= 256) rc = launch_eig<8, 8>(a, st);
  else if (n <= 384) rc = launch_eig<12, 16>(a, st);
  else rc = launch_eig<16, 16>(a, st);
  SG_CHECK(rc >= 0, "laplacian_eig: no launchable configuration for N=%d", N);
  return rc;
}

// mul_L[k] = U p_k(Lambda) U^T for k = 1..3 (p = lambda, 2 lambda^2, 4 lambda^3 - lambda; mul_L[0] = 0, base_model.py:129):
// three N x N x n products on the fp32 FFMA2 GEMM.  scratch: N*n floats.
int eig_poly_stack(const float* lambda, const float* U, int N, float* scratch, float* mul_L, cudaStream_t st) {
  const int n = (N + 1) & ~1;
  const size_t nn = (size_t)N * N;
  SG_CUDA(cudaMemsetAsync(mul_L, 0, nn * sizeof(float), st));
  for (int term = 1; term <= 3; ++term) {
    const long long total = (long long)N * n;
    scale_columns_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(U, lambda, n, N, term, scratch);
    SG_LAUNCH_CHECK("scale_columns_kernel");
    GemmOperands g = {scratch, n, 0, U, n, 0, nullptr, N, N, n};
    EpiAxpby epi = {mul_L + (size_t)term * nn, N, 0, nullptr, 0, 0, 1.f, 0.f};
    SG_TRY((launch_sgemm<false, true, false>(g, epi, 1, st, "eig_poly")));
  }
  return 0;
}

}  // namespace sg

using namespace sg;

extern "C" int stemgnn_laplacian_eig_forward(const float* attention_raw, const float* degree, int N, float* eigenvalues,
                                             float* eigenvectors, int* info, int max_sweeps, float tol,
                                             stemgnn_stream_t stream) {
  clear_error();
  SG_CHECK(attention_raw && degree && eigenvalues && eigenvectors && info, "laplacian_eig: null argument");
  return laplacian_eig(attention_raw, degree, N, eigenvalues, eigenvectors, info, max_sweeps > 0 ? max_sweeps : 30,
                       tol > 0.f ? tol : 1e-6f, static_cast<cudaStream_t>(stream));
}
