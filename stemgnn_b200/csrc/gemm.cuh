// fp32 GEMM on the sm_100a packed-FMA pipe (FFMA2 = fma.rn.f32x2, 2 fp32 FMAs per lane per issue).
//
// This is the exact-fp32 workhorse of the hot path: the graph-Fourier contraction mul_L @ x
// (base_model.py:62-64, precision-sensitive — SURVEY.md §7 hard part 3), the Chebyshev products
// (base_model.py:130-132), the folded output map, every backward GEMM, and the GLU chain when the
// tcgen05 TF32 path is disabled.  C(m,n) = sum_k A(m,k) B(k,n), optionally with a second B
// operand sharing the A tile (GLU: left and right Linear of base_model.py:12-13 in one pass).
//
// Tiling: 128x64x16 CTA tile, 256 threads, 8x4 register tile per thread, double-buffered shared
// memory with register prefetch.  Epilogues are functors (bias / GLU gate / axpby / scatter).
#pragma once
#include "common.cuh"

namespace sg {

constexpr int GM_BM = 128, GM_BN = 64, GM_BK = 16, GM_THREADS = 256;
constexpr int GM_AS_LD = GM_BM + 4, GM_BS_LD = GM_BN + 4;

struct GemmOperands {
  const float* A; int lda; long long sA;    // batch (blockIdx.z) strides in elements
  const float* B; int ldb; long long sB;
  const float* B2;                          // second B operand (DUAL), same ldb / sB
  int M, N, K;
  int ksplit;                               // > 1: blockIdx.z splits K (no batching); epilogue must be atomic
};

__device__ __forceinline__ float4 ld4_guard(const float* base, long long off, int valid, bool vec) {
  // loads up to 4 consecutive floats starting at base[off]; `valid` in [0,4] of them exist
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (valid >= 4 && vec) {
    r = __ldg(reinterpret_cast<const float4*>(base + off));
  } else {
    if (valid > 0) r.x = __ldg(base + off);
    if (valid > 1) r.y = __ldg(base + off + 1);
    if (valid > 2) r.z = __ldg(base + off + 2);
    if (valid > 3) r.w = __ldg(base + off + 3);
  }
  return r;
}

template <bool A_KM, bool B_NK, bool DUAL, class Epi>
__global__ void __launch_bounds__(GM_THREADS) sgemm_kernel(GemmOperands g, Epi epi) {
  __shared__ __align__(16) float As[2][GM_BK][GM_AS_LD];
  __shared__ __align__(16) float Bs[2][GM_BK][GM_BS_LD];
  __shared__ __align__(16) float Bs2[DUAL ? 2 : 1][DUAL ? GM_BK : 1][DUAL ? GM_BS_LD : 4];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const bool split = g.ksplit > 1;
  const int z = split ? 0 : blockIdx.z;
  int k_lo = 0, k_hi = g.K;
  if (split) {
    const int chunk = ((g.K + g.ksplit - 1) / g.ksplit + GM_BK - 1) / GM_BK * GM_BK;
    k_lo = blockIdx.z * chunk;
    k_hi = min(g.K, k_lo + chunk);
    if (k_lo >= k_hi) return;
  }
  const int m0 = blockIdx.y * GM_BM, n0 = blockIdx.x * GM_BN;
  const float* __restrict__ A = g.A + (long long)z * g.sA;
  const float* __restrict__ B = g.B + (long long)z * g.sB;
  const float* __restrict__ B2 = DUAL ? g.B2 + (long long)z * g.sB : nullptr;
  const bool vecA = ((g.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  const bool vecB = ((g.ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0) &&
                    (!DUAL || (reinterpret_cast<uintptr_t>(B2) & 15) == 0);

  float4 ra[2], rb, rb2;
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int gid = tid + i * GM_THREADS;
      if (!A_KM) {
        const int row = gid >> 2, kq = gid & 3;
        const int m = m0 + row, k = k0 + kq * 4;
        const int valid = (m < g.M) ? max(0, min(4, k_hi - k)) : 0;
        ra[i] = ld4_guard(A, (long long)m * g.lda + k, valid, vecA);
      } else {
        const int kk = gid >> 5, mq = gid & 31;
        const int k = k0 + kk, m = m0 + mq * 4;
        const int valid = (k < k_hi) ? max(0, min(4, g.M - m)) : 0;
        ra[i] = ld4_guard(A, (long long)k * g.lda + m, valid, vecA);
      }
    }
    if (B_NK) {
      const int col = tid >> 2, kq = tid & 3;
      const int n = n0 + col, k = k0 + kq * 4;
      const int valid = (n < g.N) ? max(0, min(4, k_hi - k)) : 0;
      rb = ld4_guard(B, (long long)n * g.ldb + k, valid, vecB);
      if (DUAL) rb2 = ld4_guard(B2, (long long)n * g.ldb + k, valid, vecB);
    } else {
      const int kk = tid >> 4, nq = tid & 15;
      const int k = k0 + kk, n = n0 + nq * 4;
      const int valid = (k < k_hi) ? max(0, min(4, g.N - n)) : 0;
      rb = ld4_guard(B, (long long)k * g.ldb + n, valid, vecB);
      if (DUAL) rb2 = ld4_guard(B2, (long long)k * g.ldb + n, valid, vecB);
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int gid = tid + i * GM_THREADS;
      if (!A_KM) {
        const int row = gid >> 2, kq = gid & 3;
        As[buf][kq * 4 + 0][row] = ra[i].x;
        As[buf][kq * 4 + 1][row] = ra[i].y;
        As[buf][kq * 4 + 2][row] = ra[i].z;
        As[buf][kq * 4 + 3][row] = ra[i].w;
      } else {
        const int kk = gid >> 5, mq = gid & 31;
        *reinterpret_cast<float4*>(&As[buf][kk][mq * 4]) = ra[i];
      }
    }
    if (B_NK) {
      const int col = tid >> 2, kq = tid & 3;
      Bs[buf][kq * 4 + 0][col] = rb.x;
      Bs[buf][kq * 4 + 1][col] = rb.y;
      Bs[buf][kq * 4 + 2][col] = rb.z;
      Bs[buf][kq * 4 + 3][col] = rb.w;
      if (DUAL) {
        Bs2[buf][kq * 4 + 0][col] = rb2.x;
        Bs2[buf][kq * 4 + 1][col] = rb2.y;
        Bs2[buf][kq * 4 + 2][col] = rb2.z;
        Bs2[buf][kq * 4 + 3][col] = rb2.w;
      }
    } else {
      const int kk = tid >> 4, nq = tid & 15;
      *reinterpret_cast<float4*>(&Bs[buf][kk][nq * 4]) = rb;
      if (DUAL) *reinterpret_cast<float4*>(&Bs2[buf][kk][nq * 4]) = rb2;
    }
  };

  float2 acc[8][2], acc2[DUAL ? 8 : 1][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc[i][0] = acc[i][1] = make_float2(0.f, 0.f);
    if (DUAL) acc2[i][0] = acc2[i][1] = make_float2(0.f, 0.f);
  }

  const int nk = (k_hi - k_lo + GM_BK - 1) / GM_BK;
  load_tiles(k_lo);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles(k_lo + (kt + 1) * GM_BK);
#pragma unroll
    for (int kk = 0; kk < GM_BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float2 b01 = make_float2(b.x, b.y), b23 = make_float2(b.z, b.w);
      float2 c01, c23;
      if (DUAL) {
        const float4 bb = *reinterpret_cast<const float4*>(&Bs2[buf][kk][tx * 4]);
        c01 = make_float2(bb.x, bb.y);
        c23 = make_float2(bb.z, bb.w);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 aa = make_float2(av[i], av[i]);
        acc[i][0] = __ffma2_rn(aa, b01, acc[i][0]);
        acc[i][1] = __ffma2_rn(aa, b23, acc[i][1]);
        if (DUAL) {
          acc2[i][0] = __ffma2_rn(aa, c01, acc2[i][0]);
          acc2[i][1] = __ffma2_rn(aa, c23, acc2[i][1]);
        }
      }
    }
    if (kt + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

  const int n = n0 + tx * 4;
  if (n < g.N) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = m0 + ty * 8 + i;
      if (m >= g.M) break;
      const float4 v = make_float4(acc[i][0].x, acc[i][0].y, acc[i][1].x, acc[i][1].y);
      float4 v2 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (DUAL) v2 = make_float4(acc2[i][0].x, acc2[i][0].y, acc2[i][1].x, acc2[i][1].y);
      epi.store4(z, m, n, min(4, g.N - n), v, v2);
    }
  }
}

// ---- epilogues ----------------------------------------------------------------------------
__device__ __forceinline__ void st4_guard(float* base, long long off, int valid, float4 v) {
  if (valid >= 4 && ((off & 3) == 0) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0)) {
    *reinterpret_cast<float4*>(base + off) = v;
  } else {
    if (valid > 0) base[off] = v.x;
    if (valid > 1) base[off + 1] = v.y;
    if (valid > 2) base[off + 2] = v.z;
    if (valid > 3) base[off + 3] = v.w;
  }
}

// C = alpha * AB + beta * Cin  (Cin may be null when beta == 0); optional row-bias / col-bias none
struct EpiAxpby {
  float* C; int ldc; long long sC;
  const float* Cin; int ldcin; long long sCin;
  float alpha, beta;
  __device__ __forceinline__ void store4(int z, int m, int n, int valid, float4 v, float4) const {
    float4 o = make_float4(alpha * v.x, alpha * v.y, alpha * v.z, alpha * v.w);
    if (Cin != nullptr) {
      const float* ci = Cin + (long long)z * sCin + (long long)m * ldcin + n;
      if (valid > 0) o.x += beta * ci[0];
      if (valid > 1) o.y += beta * ci[1];
      if (valid > 2) o.z += beta * ci[2];
      if (valid > 3) o.w += beta * ci[3];
    }
    st4_guard(C + (long long)z * sC, (long long)m * ldc + n, valid, o);
  }
};

// C = AB + bias[n]   (ACT: 0 none, 1 sigmoid)
template <int ACT>
struct EpiBias {
  float* C; int ldc; long long sC;
  const float* bias; long long sBias;      // may be null
  __device__ __forceinline__ void store4(int z, int m, int n, int valid, float4 v, float4) const {
    float o[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < valid) {
        if (bias != nullptr) o[j] += __ldg(bias + (long long)z * sBias + n + j);
        if (ACT == 1) o[j] = sigmoidf_(o[j]);
      }
    }
    st4_guard(C + (long long)z * sC, (long long)m * ldc + n, valid, make_float4(o[0], o[1], o[2], o[3]));
  }
};

// GLU gate (base_model.py:12-13): out = (AB_l + b_l) * sigmoid(AB_r + b_r).
// When `save_l` / `save_s` are non-null (training) the left pre-activation and the gate value are
// kept for the backward pass.
struct EpiGlu {
  float* out; int ldo; long long sO;
  const float* bl; const float* br; long long sBias;
  float* save_l; float* save_s; int lds; long long sS;
  __device__ __forceinline__ void store4(int z, int m, int n, int valid, float4 v, float4 v2) const {
    float l[4] = {v.x, v.y, v.z, v.w}, r[4] = {v2.x, v2.y, v2.z, v2.w}, o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (j < valid) {
        l[j] += __ldg(bl + (long long)z * sBias + n + j);
        r[j] = sigmoidf_(r[j] + __ldg(br + (long long)z * sBias + n + j));
        o[j] = l[j] * r[j];
      } else {
        o[j] = 0.f;
      }
    }
    st4_guard(out + (long long)z * sO, (long long)m * ldo + n, valid, make_float4(o[0], o[1], o[2], o[3]));
    if (save_l != nullptr) {
      st4_guard(save_l + (long long)z * sS, (long long)m * lds + n, valid, make_float4(l[0], l[1], l[2], l[3]));
      st4_guard(save_s + (long long)z * sS, (long long)m * lds + n, valid, make_float4(r[0], r[1], r[2], r[3]));
    }
  }
};

// C += alpha * AB with atomics (split-K partial sums; gradients accumulate by contract)
struct EpiAtomicAdd {
  float* C; int ldc;
  float alpha;
  __device__ __forceinline__ void store4(int, int m, int n, int valid, float4 v, float4) const {
    float* c = C + (long long)m * ldc + n;
    if (valid > 0) atomicAdd(c, alpha * v.x);
    if (valid > 1) atomicAdd(c + 1, alpha * v.y);
    if (valid > 2) atomicAdd(c + 2, alpha * v.z);
    if (valid > 3) atomicAdd(c + 3, alpha * v.w);
  }
};

// deterministic split-K: split z writes its partial product to P[z] (M x ldp); a fixed-order reduction
// kernel combines them afterwards (bitwise reproducible, unlike atomics)
struct EpiPartial {
  float* P; int ldp; long long stride;
  __device__ __forceinline__ void store4(int, int m, int n, int valid, float4 v, float4) const {
    st4_guard(P + (long long)blockIdx.z * stride, (long long)m * ldp + n, valid, v);
  }
};

// number of K-splits that brings a (M x N x K) product to about two waves of CTAs
inline int pick_ksplit(int M, int N, int K) {
  const int tiles = ceil_div(M, GM_BM) * ceil_div(N, GM_BN);
  int s = (2 * 148 + tiles - 1) / tiles;
  const int maxs = K / (4 * GM_BK) > 0 ? K / (4 * GM_BK) : 1;
  if (s > maxs) s = maxs;
  return s < 1 ? 1 : s;
}

template <bool A_KM, bool B_NK, bool DUAL, class Epi>
inline int launch_sgemm(const GemmOperands& g, const Epi& epi, int batch, cudaStream_t st,
                        const char* name) {
  if (g.M <= 0 || g.N <= 0 || batch <= 0) return 0;
  if (g.ksplit > 1) batch = g.ksplit;
  dim3 grid(ceil_div(g.N, GM_BN), ceil_div(g.M, GM_BM), batch);
  sgemm_kernel<A_KM, B_NK, DUAL, Epi><<<grid, GM_THREADS, 0, st>>>(g, epi);
  SG_LAUNCH_CHECK(name);
  return 0;
}

}  // namespace sg
