// GRU over the node axis + key/query contraction  (reference: models/base_model.py:92,137 and
// :154-155; gate equations as in torch.nn.GRU, gate order [r,z,n], h0 = 0).
//
// The recurrence is N dependent steps of h(B x N) . W_hh^T(N x 3N): latency-bound, O(B N^3).
// B200 design (DESIGN.md "GRU"):
//   * batch elements are independent sequences -> one THREAD-BLOCK CLUSTER of CS=16 CTAs per
//     group of 4 batch elements (8 clusters = 128 SMs at B=32);
//   * each CTA of a cluster owns ceil(N/16) hidden units: its 3*U rows of W_hh stay resident in
//     shared memory for all N steps (persistent-RNN), so W_hh is read from HBM exactly once;
//   * per step: packed-fp32 (FFMA2) mat-vec against the 4 hidden vectors, gate math, then the new
//     hidden slice is scattered into every CTA's next-step buffer through DISTRIBUTED SHARED
//     MEMORY and one cluster barrier (arrive / wait split; the input projection W_ih x_{s+1} of
//     the next step is computed between arrive and wait);
//   * key/query (sum over steps of h_s * w[s]) are accumulated in registers, so the (N,B,N) GRU
//     output is never materialised in eval mode (it is written only when the backward needs it).
// A generic per-step-launch path covers N > 512 or devices that refuse the 16-CTA cluster.
#include <cooperative_groups.h>

#include "common.cuh"
#include "internal.cuh"

namespace cg = cooperative_groups;

namespace sg {


constexpr int GRU_BC = 4;        // batch elements per cluster
constexpr int GRU_WARPS = 8;
constexpr int GRU_THREADS = GRU_WARPS * 32;

__device__ __forceinline__ void cluster_arrive_release() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void cluster_wait_acquire() {
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// JC: K padded to 128*JC;  UPW: hidden units per warp;  CS: cluster size
template <int JC, int UPW, int CS>
__global__ void __launch_bounds__(GRU_THREADS, 1) gru_cluster_kernel(GruArgs a) {
  constexpr int KP = 128 * JC;
  constexpr int NJ = KP / 32;               // float4 chunks per lane
  constexpr int ULOC = GRU_WARPS * UPW;     // padded units per CTA
  constexpr int ROWS = 3 * UPW;             // rows per warp

  extern __shared__ __align__(16) float smem[];
  float* Wsm = smem;                              // [3*ULOC][KP]
  float* hbuf = Wsm + 3 * ULOC * KP;              // [2][BC][KP]
  float* stage = hbuf + 2 * GRU_BC * KP;          // [BC][32]

  cg::cluster_group cluster = cg::this_cluster();
  const int q = (int)cluster.block_rank();
  const int cid = blockIdx.x / CS;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int b = lane >> 3, kg = lane & 7;
  const int N = a.N, B = a.B, W = a.W;
  const int U = (N + CS - 1) / CS;
  const int u0 = q * U;
  const int b0 = cid * GRU_BC;

  // ---- one-time: W_hh slice -> smem (zero padded), h buffers = 0 ---------------------------
  for (int idx = tid; idx < 3 * ULOC * KP; idx += GRU_THREADS) {
    const int row = idx / KP, k = idx - row * KP;
    const int lu = row / 3, gate = row - lu * 3;
    const int u = u0 + lu;
    float v = 0.f;
    if (lu < U && u < N && k < N) v = __ldg(a.w_hh + ((long long)gate * N + u) * N + k);
    Wsm[idx] = v;
  }
  for (int idx = tid; idx < 2 * GRU_BC * KP; idx += GRU_THREADS) hbuf[idx] = 0.f;

  // finalising lanes: lane (b, kg) with kg < UPW owns local unit lu = w*UPW + kg for batch b
  const int lu = w * UPW + kg;
  const int u = u0 + lu;
  const bool fin = (kg < UPW) && (lu < U) && (u < N);
  const bool bvalid = (b0 + b) < B;
  float bhr = 0.f, bhz = 0.f, bhn = 0.f;
  if (fin) {
    bhr = __ldg(a.b_hh + u);
    bhz = __ldg(a.b_hh + N + u);
    bhn = __ldg(a.b_hh + 2 * N + u);
  }
  float key_acc = 0.f, query_acc = 0.f;

  // input projection of step s for this lane's (unit, batch): W_i{r,z,n} x_s + b_i{r,z,n}
  auto input_proj = [&](int s, float& gr, float& gz, float& gn) {
    gr = gz = gn = 0.f;
    if (fin) {
      gr = __ldg(a.b_ih + u);
      gz = __ldg(a.b_ih + N + u);
      gn = __ldg(a.b_ih + 2 * N + u);
      if (bvalid) {
        const float* xrow = a.xs + ((long long)s * B + (b0 + b)) * W;
        const float* wr = a.w_ih + (long long)u * W;
        const float* wz = a.w_ih + (long long)(N + u) * W;
        const float* wn = a.w_ih + (long long)(2 * N + u) * W;
        for (int t = 0; t < W; ++t) {
          const float xv = __ldg(xrow + t);
          gr = fmaf(__ldg(wr + t), xv, gr);
          gz = fmaf(__ldg(wz + t), xv, gz);
          gn = fmaf(__ldg(wn + t), xv, gn);
        }
      }
    }
  };

  float gi_r, gi_z, gi_n;
  input_proj(0, gi_r, gi_z, gi_n);
  float wk_s = __ldg(a.wk + 0), wq_s = __ldg(a.wq + 0);

  __syncthreads();
  cluster.sync();   // every CTA's buffers are initialised before any remote write

  for (int s = 0; s < N; ++s) {
    const int cur = s & 1, nxt = cur ^ 1;
    // (a) this lane's slice of h_{s-1} for batch b: k = 32*jj + 4*kg + {0..3}
    const float* hb = hbuf + (cur * GRU_BC + b) * KP + 4 * kg;
    float4 h[NJ];
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) h[jj] = *reinterpret_cast<const float4*>(hb + 32 * jj);

    // (b) mat-vec: ROWS rows of this warp against the lane's k-slice (two partial sums / row)
    float2 acc[ROWS];
    const float* wbase = Wsm + (long long)(w * ROWS) * KP + 4 * kg;
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      acc[r] = make_float2(0.f, 0.f);
      const float* wrow = wbase + r * KP;
#pragma unroll
      for (int jj = 0; jj < NJ; ++jj) {
        const float4 wv = *reinterpret_cast<const float4*>(wrow + 32 * jj);
        acc[r] = __ffma2_rn(make_float2(wv.x, wv.y), make_float2(h[jj].x, h[jj].y), acc[r]);
        acc[r] = __ffma2_rn(make_float2(wv.z, wv.w), make_float2(h[jj].z, h[jj].w), acc[r]);
      }
    }
    // (c) reduce over the 8 k-groups (lanes sharing b)
    float sum[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      float v = acc[r].x + acc[r].y;
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      sum[r] = v;
    }
    // (d) gates for (unit lu, batch b) on lanes kg < UPW
    float gh_r = 0.f, gh_z = 0.f, gh_n = 0.f;
#pragma unroll
    for (int i = 0; i < UPW; ++i) {
      if (kg == i) {
        gh_r = sum[3 * i];
        gh_z = sum[3 * i + 1];
        gh_n = sum[3 * i + 2];
      }
    }
    if (kg < UPW) {
      float hn = 0.f;
      if (fin) {
        const float hprev = hbuf[(cur * GRU_BC + b) * KP + u];
        const float r = sigmoidf_(gi_r + gh_r + bhr);
        const float zt = sigmoidf_(gi_z + gh_z + bhz);
        const float nt = tanhf(gi_n + r * (gh_n + bhn));
        hn = (1.f - zt) * nt + zt * hprev;
        key_acc = fmaf(hn, wk_s, key_acc);
        query_acc = fmaf(hn, wq_s, query_acc);
      }
      stage[b * 32 + lu] = hn;
    }
    __syncthreads();
    // (e) scatter the CTA's new slice into every cluster CTA's next buffer (DSMEM)
    if (lane < U && (u0 + lane) < N) {
#pragma unroll
      for (int d = w; d < CS; d += GRU_WARPS) {
        float* remote = cluster.map_shared_rank(hbuf, d);
#pragma unroll
        for (int bb = 0; bb < GRU_BC; ++bb)
          remote[(nxt * GRU_BC + bb) * KP + u0 + lane] = stage[bb * 32 + lane];
      }
    }
    if (a.h_all != nullptr && w < GRU_BC && (b0 + w) < B && lane < U && (u0 + lane) < N)
      a.h_all[((long long)s * B + (b0 + w)) * N + u0 + lane] = stage[w * 32 + lane];
    // (f) barrier; overlap the next step's input projection with the wait
    __syncwarp();
    cluster_arrive_release();
    if (s + 1 < N) {
      input_proj(s + 1, gi_r, gi_z, gi_n);
      wk_s = __ldg(a.wk + s + 1);
      wq_s = __ldg(a.wq + s + 1);
    }
    __syncwarp();
    cluster_wait_acquire();
  }

  if (fin && bvalid) {
    a.key[(long long)(b0 + b) * N + u] = key_acc;
    a.query[(long long)(b0 + b) * N + u] = query_acc;
  }
}

template <int JC, int UPW, int CS>
static int launch_gru_cluster(const GruArgs& a, cudaStream_t st, bool probe_only) {
  constexpr int KP = 128 * JC;
  constexpr int ULOC = GRU_WARPS * UPW;
  const size_t smem = (size_t)(3 * ULOC * KP + 2 * GRU_BC * KP + GRU_BC * 32) * sizeof(float);
  auto kern = gru_cluster_kernel<JC, UPW, CS>;
  static bool attr_set = false;   // one process drives one device (DDP = process per GPU)
  if (!attr_set) {
    SG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (CS > 8)
      SG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    attr_set = true;
  }
  const int nclusters = ceil_div(a.B, GRU_BC);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(nclusters * CS);
  cfg.blockDim = dim3(GRU_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int max_clusters = 0;
  cudaError_t e = cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg);
  if (e != cudaSuccess || max_clusters < 1) {
    (void)cudaGetLastError();
    return -1;   // not launchable with this cluster size on this device
  }
  if (probe_only) return 0;
  SG_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
  return 0;
}

template <int CS>
static int dispatch_gru_cluster(const GruArgs& a, cudaStream_t st) {
  const int U = ceil_div(a.N, CS);
  const int upw = ceil_div(U, GRU_WARPS);
  const int jc = ceil_div(a.N, 128);
  if (upw > 4 || jc > 4) return -1;
#define SG_GRU_CASE(J, P) \
  if (jc == J && upw == P) return launch_gru_cluster<J, P, CS>(a, st, false);
  SG_GRU_CASE(1, 1) SG_GRU_CASE(1, 2) SG_GRU_CASE(1, 3) SG_GRU_CASE(1, 4)
  SG_GRU_CASE(2, 1) SG_GRU_CASE(2, 2) SG_GRU_CASE(2, 3) SG_GRU_CASE(2, 4)
  SG_GRU_CASE(3, 1) SG_GRU_CASE(3, 2) SG_GRU_CASE(3, 3) SG_GRU_CASE(3, 4)
  SG_GRU_CASE(4, 1) SG_GRU_CASE(4, 2) SG_GRU_CASE(4, 3) SG_GRU_CASE(4, 4)
#undef SG_GRU_CASE
  return -1;
}

// ---- generic path: one launch per step, W_hh streamed from L2 --------------------------------
// grid.x = ceil(N / 4) (one warp per hidden unit), loops over the batch in chunks of 4.
__global__ void __launch_bounds__(128) gru_step_kernel(GruArgs a, int s, const float* __restrict__ h_prev,
                                                       float* __restrict__ h_next) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int u = blockIdx.x * 4 + wid;
  const int N = a.N, B = a.B, W = a.W;
  if (u >= N) return;
  const float* wr = a.w_hh + (long long)u * N;
  const float* wz = a.w_hh + (long long)(N + u) * N;
  const float* wn = a.w_hh + (long long)(2 * N + u) * N;
  const float wk_s = __ldg(a.wk + s), wq_s = __ldg(a.wq + s);
  for (int bb = 0; bb < B; bb += 4) {
    float acc[3][4];
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[g][j] = 0.f;
    for (int k = lane; k < N; k += 32) {
      const float r = __ldg(wr + k), z = __ldg(wz + k), n = __ldg(wn + k);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float hv = (bb + j < B) ? h_prev[(long long)(bb + j) * N + k] : 0.f;
        acc[0][j] = fmaf(r, hv, acc[0][j]);
        acc[1][j] = fmaf(z, hv, acc[1][j]);
        acc[2][j] = fmaf(n, hv, acc[2][j]);
      }
    }
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[g][j] = warp_sum(acc[g][j]);
    if (lane < 4 && bb + lane < B) {
      const int bi = bb + lane;
      float gh_r = 0.f, gh_z = 0.f, gh_n = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (lane == j) { gh_r = acc[0][j]; gh_z = acc[1][j]; gh_n = acc[2][j]; }
      float gr = __ldg(a.b_ih + u), gz = __ldg(a.b_ih + N + u), gn = __ldg(a.b_ih + 2 * N + u);
      const float* xrow = a.xs + ((long long)s * B + bi) * W;
      for (int t = 0; t < W; ++t) {
        const float xv = __ldg(xrow + t);
        gr = fmaf(__ldg(a.w_ih + (long long)u * W + t), xv, gr);
        gz = fmaf(__ldg(a.w_ih + (long long)(N + u) * W + t), xv, gz);
        gn = fmaf(__ldg(a.w_ih + (long long)(2 * N + u) * W + t), xv, gn);
      }
      const float r = sigmoidf_(gr + gh_r + __ldg(a.b_hh + u));
      const float zt = sigmoidf_(gz + gh_z + __ldg(a.b_hh + N + u));
      const float nt = tanhf(gn + r * (gh_n + __ldg(a.b_hh + 2 * N + u)));
      const float hn = (1.f - zt) * nt + zt * h_prev[(long long)bi * N + u];
      h_next[(long long)bi * N + u] = hn;
      if (a.h_all != nullptr) a.h_all[((long long)s * B + bi) * N + u] = hn;
      // key/query accumulate in place (this thread is the only writer of (bi,u))
      a.key[(long long)bi * N + u] = fmaf(hn, wk_s, a.key[(long long)bi * N + u]);
      a.query[(long long)bi * N + u] = fmaf(hn, wq_s, a.query[(long long)bi * N + u]);
    }
  }
}

// scratch: 2*B*N floats (ping-pong hidden state) for the generic path
int gru_keyquery_forward(const GruArgs& a, int path, float* scratch, cudaStream_t st) {
  SG_CHECK(a.B > 0 && a.N > 0 && a.W > 0, "gru: bad dims B=%d N=%d W=%d", a.B, a.N, a.W);
  if (path != 1) {
    int rc = dispatch_gru_cluster<16>(a, st);
    if (rc == 0) return 0;
    if (rc > 0) return rc;
    if (a.N <= 256) {
      rc = dispatch_gru_cluster<8>(a, st);
      if (rc == 0) return 0;
      if (rc > 0) return rc;
    }
    SG_CHECK(path != 2, "gru: cluster path unavailable for N=%d on this device", a.N);
  }
  const size_t bn = (size_t)a.B * a.N;
  SG_CUDA(cudaMemsetAsync(scratch, 0, 2 * bn * sizeof(float), st));
  SG_CUDA(cudaMemsetAsync(a.key, 0, bn * sizeof(float), st));
  SG_CUDA(cudaMemsetAsync(a.query, 0, bn * sizeof(float), st));
  for (int s = 0; s < a.N; ++s) {
    const float* hp = scratch + (size_t)(s & 1) * bn;
    float* hn = scratch + (size_t)((s & 1) ^ 1) * bn;
    gru_step_kernel<<<ceil_div(a.N, 4), 128, 0, st>>>(a, s, hp, hn);
  }
  SG_LAUNCH_CHECK("gru_step_kernel");
  return 0;
}

}  // namespace sg
