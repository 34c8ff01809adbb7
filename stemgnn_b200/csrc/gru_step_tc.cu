// GRU recurrence for node counts beyond the cluster envelope (N > 480: BASELINE.json configs[4], N = 2048) on tcgen05.
// Reference: models/base_model.py:92,137 (nn.GRU over the node axis) and :154-155 (key / query contraction).
//
// W_hh (3N x N) no longer fits on chip (fp16 hi + lo = 50 MB at N = 2048), so a step is ONE kernel launch whose CTAs
// stream their rows of W_hh from L2 (126 MB: the hi/lo images stay L2-resident across the N steps) through TMA:
//   * CTA (q, grp) owns 40 hidden units (120 rows of W_hh, gate-major) and 32 sequences; per 64-column K chunk a TMA stage
//     carries the A tiles W_hi, W_lo (128 x 64 fp16 each) and the B tile [h_hi (32 rows); h_lo (32 rows)] of h_{s-1};
//   * two M128 x N64 x K16 instructions per 16 hidden units (A = W_hi, A = W_lo) accumulate the four partial products in
//     tensor memory; gh = D00 + 2^-11 (D01 + D10) + 2^-22 D11 (same split-operand scheme as gru_tc.cu: fp32-level parity);
//   * epilogue: tcgen05.ld -> smem transpose -> gate math with the in-kernel input projection (36 FMAs per unit-step) ->
//     fp32 master copy of h, key / query accumulators (unit-major scratch) and the fp16 hi/lo image of h_s that the NEXT
//     launch reads as its B operand (kernel boundary = the grid-wide exchange of h).
// Bound: L2 -> SM streaming of the W_hh images (2 * 3N * N * 2 bytes per step = 50 MB at N = 2048, ~10 us at the measured
// ~5-6 TB/s L2 rate) — DESIGN.md §8.  The round-1 per-step FFMA path (gru_step_kernel) took ~130 us per step.
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "common.cuh"
#include "internal.cuh"

namespace sg {
namespace {

constexpr int ST_U = 40;                // hidden units per CTA
constexpr int ST_SEQ = 32;              // sequences per CTA (B operand: 32 hi rows + 32 lo rows)
constexpr int ST_THREADS = 192;         // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue
constexpr int ST_NSTG = 4;
constexpr uint32_t ST_A_BYTES = 128 * 128;        // one A tile (128 rows x 64 fp16)
constexpr uint32_t ST_B_BYTES = 64 * 128;         // B tile (64 rows x 64 fp16)
constexpr uint32_t ST_STAGE = 2 * ST_A_BYTES + ST_B_BYTES;
constexpr uint32_t ST_TMEM_COLS = 128;
constexpr float ST_LO_SCALE = 2048.0f, ST_LO_INV = 1.0f / 2048.0f, ST_LO_INV2 = 1.0f / (2048.0f * 2048.0f);

__device__ __forceinline__ uint32_t sa(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbi(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sa(bar)), "r"(count));
}
__device__ __forceinline__ void mbx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sa(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbw(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = sa(bar);
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (spin > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void tma2(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(sa(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(sa(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tfb() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tfa() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint32_t el1() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred;
}
__device__ __forceinline__ uint64_t dsc128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
constexpr uint32_t ST_IDESC = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);     // fp16 x fp16 -> fp32, M128 N64
__device__ __forceinline__ void mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(ST_IDESC), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void cmt(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sa(bar)) : "memory");
}
__device__ __forceinline__ void tld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tldw() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float fsig(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float ftanh(float x) { return 2.0f * fsig(2.0f * x) - 1.0f; }
__device__ __forceinline__ void split16(float x, unsigned short& hi, unsigned short& lo) {
  const __half h = __float2half_rn(x);
  hi = __half_as_ushort(h);
  lo = __half_as_ushort(__float2half_rn((x - __half2float(h)) * ST_LO_SCALE));
}

// W_hh (3N, N) fp32 -> image rows [(tile * 2 + arr) * 128 + (gate * 40 + lu)][KP] fp16 (rows 120..127 and pads zero)
__global__ void __launch_bounds__(256) step_pack_whh_kernel(const float* __restrict__ w_hh, unsigned short* __restrict__ img,
                                                             int N, int NT, int KP) {
  const long long total = (long long)NT * 128 * KP;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(idx % KP);
    const int row = (int)((idx / KP) % 128);
    const int q = (int)(idx / ((long long)KP * 128));
    float v = 0.f;
    if (row < 3 * ST_U && k < N) {
      const int gate = row / ST_U, lu = row - gate * ST_U;
      const int u = q * ST_U + lu;
      if (u < N) v = __ldg(w_hh + ((size_t)gate * N + u) * N + k);
    }
    unsigned short hi, lo;
    split16(v, hi, lo);
    img[((size_t)(q * 2 + 0) * 128 + row) * KP + k] = hi;
    img[((size_t)(q * 2 + 1) * 128 + row) * KP + k] = lo;
  }
}

struct StepArgs {
  GruArgs a;
  float* hstate;          // (N, Bp) fp32 master copy of h, unit-major (Bp = groups * 32)
  float* keyT;            // (N, Bp)
  float* queryT;          // (N, Bp)
  unsigned short* himg;   // [2][groups * 64][KP] fp16 images of h (rows 0..31 hi, 32..63 lo per group)
  int s, NCH, KP, Bp;
};

__global__ void __launch_bounds__(ST_THREADS, 1)
gru_step_tc_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_h, StepArgs t) {
  extern __shared__ uint8_t smem_raw[];
  // aligned by pointer arithmetic on the __shared__ array (an integer round trip demotes every access to generic LD/ST)
  uint8_t* smem = smem_raw + ((1024u - (sa(smem_raw) & 1023u)) & 1023u);
  uint8_t* stages = smem;
  float* gh_sm = reinterpret_cast<float*>(stages + ST_NSTG * ST_STAGE);     // [120][32]
  float* wih_sm = gh_sm + 3 * ST_U * ST_SEQ;                                 // [120][W]
  float* bias_sm = wih_sm + 3 * ST_U * t.a.W;                                // [4][40]
  float* x_sm = bias_sm + 4 * ST_U;                                          // [32][W]
  uint8_t* bar_base = reinterpret_cast<uint8_t*>(x_sm + ST_SEQ * t.a.W);
  bar_base += (8u - (sa(bar_base) & 7u)) & 7u;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(bar_base);
  uint64_t* empty_bar = full_bar + ST_NSTG;
  uint64_t* tfull = empty_bar + ST_NSTG;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const GruArgs& a = t.a;
  const int N = a.N, B = a.B, W = a.W, s = t.s;
  const int q = blockIdx.x, grp = blockIdx.y;
  const int u0 = q * ST_U, b0 = grp * ST_SEQ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cur = s & 1, nxt = cur ^ 1;

  if (threadIdx.x == 0) {
    for (int i = 0; i < ST_NSTG; ++i) {
      mbi(&full_bar[i], 1);
      mbi(&empty_bar[i], 1);
    }
    mbi(tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sa(tmem_slot)), "r"(ST_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= 2) {      // step-local constants for the gate phase
    const int et = threadIdx.x - 64;
    for (int i = et; i < 3 * ST_U * W; i += 128) {
      const int row = i / W, tt = i - row * W;
      const int gate = row / ST_U, lu = row - gate * ST_U;
      const int u = u0 + lu;
      wih_sm[i] = u < N ? __ldg(a.w_ih + ((size_t)gate * N + u) * W + tt) : 0.f;
    }
    for (int i = et; i < 4 * ST_U; i += 128) {
      const int gate = i / ST_U, lu = i - gate * ST_U;
      const int u = u0 + lu;
      float v = 0.f;
      if (u < N) {
        if (gate < 2) v = __ldg(a.b_ih + gate * N + u) + __ldg(a.b_hh + gate * N + u);
        else if (gate == 2) v = __ldg(a.b_ih + 2 * N + u);
        else v = __ldg(a.b_hh + 2 * N + u);
      }
      bias_sm[i] = v;
    }
    for (int i = et; i < ST_SEQ * W; i += 128) {
      const int bb = i / W, tt = i - bb * W;
      x_sm[i] = (b0 + bb) < B ? __ldg(a.xs + ((size_t)s * B + b0 + bb) * W + tt) : 0.f;
    }
  }
  tfb();
  __syncthreads();
  tfa();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (el1()) {        // ===== TMA producer =====
      for (int c = 0; c < t.NCH; ++c) {
        const int st = c % ST_NSTG;
        const uint32_t ph = (uint32_t)(c / ST_NSTG) & 1u;
        mbw(&empty_bar[st], ph ^ 1u);
        uint8_t* dst = stages + (size_t)st * ST_STAGE;
        mbx(&full_bar[st], ST_STAGE);
        tma2(dst, &map_w, &full_bar[st], c * 64, (q * 2 + 0) * 128);
        tma2(dst + ST_A_BYTES, &map_w, &full_bar[st], c * 64, (q * 2 + 1) * 128);
        tma2(dst + 2 * ST_A_BYTES, &map_h, &full_bar[st], c * 64, (cur * (int)gridDim.y + grp) * 64);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer (converged warp, elected lane) =====
    for (int c = 0; c < t.NCH; ++c) {
      const int st = c % ST_NSTG;
      const uint32_t ph = (uint32_t)(c / ST_NSTG) & 1u;
      mbw(&full_bar[st], ph);
      tfa();
      if (el1()) {
        const uint32_t a_hi = sa(stages + (size_t)st * ST_STAGE), a_lo = a_hi + ST_A_BYTES, bt = a_lo + ST_A_BYTES;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t bd = dsc128(bt + kk * 32);
          const uint32_t acc = (c > 0 || kk > 0) ? 1u : 0u;
          mma(tmem_base, dsc128(a_hi + kk * 32), bd, acc);            // W_hi . [h_hi | h_lo]
          mma(tmem_base + 64, dsc128(a_lo + kk * 32), bd, acc);       // W_lo . [h_hi | h_lo]
        }
        cmt(&empty_bar[st]);
        if (c + 1 == t.NCH) cmt(tfull);
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue =====
    const int et = threadIdx.x - 64;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    mbw(tfull, 0);
    tfa();
#pragma unroll
    for (int half = 0; half < 2; ++half) {            // 16 sequences at a time; every lane takes part in the aligned loads
      float d00[16], d01[16], d10[16], d11[16];
      tld16(taddr + half * 16, d00);                   // W_hi.h_hi
      tld16(taddr + 32 + half * 16, d01);              // W_hi.h_lo
      tld16(taddr + 64 + half * 16, d10);              // W_lo.h_hi
      tld16(taddr + 96 + half * 16, d11);              // W_lo.h_lo
      tldw();
      if (row < 3 * ST_U) {                            // rows 120..127 hold no unit
#pragma unroll
        for (int j = 0; j < 16; ++j)
          gh_sm[row * ST_SEQ + half * 16 + j] = d00[j] + (d01[j] + d10[j]) * ST_LO_INV + d11[j] * ST_LO_INV2;
      }
    }
    tfb();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    // gate phase: thread -> sequence b = et % 32, units lu = et / 32 + 4 i
    const int bb = et & 31;
    const int bglob = b0 + bb;
    const float wk_s = __ldg(a.wk + s), wq_s = __ldg(a.wq + s);
    unsigned short* hnext = t.himg + ((size_t)(nxt * (int)gridDim.y + grp) * 64) * t.KP;
    for (int lu = et >> 5; lu < ST_U; lu += 4) {
      const int u = u0 + lu;
      float hn = 0.f;
      if (u < N && bglob < B) {
        float gr = bias_sm[lu], gz = bias_sm[ST_U + lu], gn = bias_sm[2 * ST_U + lu];
        const float* xr = x_sm + bb * W;
        const float* wr = wih_sm + (size_t)lu * W;
        const float* wz = wr + (size_t)ST_U * W;
        const float* wn = wz + (size_t)ST_U * W;
        for (int tt = 0; tt < W; ++tt) {
          const float xv = xr[tt];
          gr = fmaf(wr[tt], xv, gr);
          gz = fmaf(wz[tt], xv, gz);
          gn = fmaf(wn[tt], xv, gn);
        }
        const float gh_r = gh_sm[lu * ST_SEQ + bb], gh_z = gh_sm[(ST_U + lu) * ST_SEQ + bb];
        const float gh_n = gh_sm[(2 * ST_U + lu) * ST_SEQ + bb] + bias_sm[3 * ST_U + lu];
        const size_t so = (size_t)u * t.Bp + bglob;
        const float hprev = s > 0 ? t.hstate[so] : 0.f;
        const float r = fsig(gr + gh_r), z = fsig(gz + gh_z), n = ftanh(gn + r * gh_n);
        hn = (1.f - z) * n + z * hprev;
        t.hstate[so] = hn;
        t.keyT[so] = fmaf(hn, wk_s, s > 0 ? t.keyT[so] : 0.f);
        t.queryT[so] = fmaf(hn, wq_s, s > 0 ? t.queryT[so] : 0.f);
        if (a.h_all != nullptr) {
          const size_t o = ((size_t)s * B + bglob) * N + u;
          a.h_all[o] = hn;
          if (a.g_r != nullptr) {
            a.g_r[o] = r; a.g_z[o] = z; a.g_n[o] = n; a.g_hn[o] = gh_n;
          }
        }
      }
      if (u < t.KP) {      // fp16 hi / lo image of h_s: the B operand of the next launch
        unsigned short hi, lo;
        split16(hn, hi, lo);
        hnext[(size_t)bb * t.KP + u] = hi;
        hnext[(size_t)(32 + bb) * t.KP + u] = lo;
      }
    }
  }
  tfb();
  __syncthreads();
  if (warp == 1) {
    tfa();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ST_TMEM_COLS) : "memory");
  }
}

// key / query (N, Bp) unit-major scratch -> (B, N)
__global__ void step_finish_kernel(const float* __restrict__ keyT, const float* __restrict__ queryT, float* __restrict__ key,
                                   float* __restrict__ query, int B, int N, int Bp) {
  const long long total = (long long)B * N;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(idx / N), u = (int)(idx % N);
    key[idx] = keyT[(size_t)u * Bp + b];
    query[idx] = queryT[(size_t)u * Bp + b];
  }
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncFn enc_fn() {
  static EncFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncFn>(p);
    else
      (void)cudaGetLastError();
  }
  return fn;
}
bool map16(EncFn enc, CUtensorMap* map, const void* base, long long rows, int cols, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

// bytes of scratch (inside the `gi` region) the step path needs
size_t gru_step_tc_scratch_bytes(int B, int N) {
  const int NT = ceil_div(N, ST_U), KP = ceil_div(N, 64) * 64, groups = ceil_div(B, ST_SEQ), Bp = groups * ST_SEQ;
  return (size_t)NT * 2 * 128 * KP * 2 + (size_t)2 * groups * 64 * KP * 2 + (size_t)3 * KP * Bp * 4 + 1024;
}

// returns 0 launched, -1 not applicable here (caller falls back), >0 error
int gru_step_tc_forward(const GruArgs& a, uint8_t* scratch, size_t scratch_bytes, int reuse_img, cudaStream_t st) {
  static const bool off = getenv("STEMGNN_GRU_NO_STEP_TC") != nullptr;
  if (off || scratch == nullptr || a.W > 64) return -1;
  if (gru_step_tc_scratch_bytes(a.B, a.N) > scratch_bytes || (reinterpret_cast<uintptr_t>(scratch) & 255)) return -1;
  EncFn enc = enc_fn();
  if (enc == nullptr) return -1;
  const int N = a.N, B = a.B;
  const int NT = ceil_div(N, ST_U), NCH = ceil_div(N, 64), KP = NCH * 64, groups = ceil_div(B, ST_SEQ), Bp = groups * ST_SEQ;
  unsigned short* wimg = reinterpret_cast<unsigned short*>(scratch);
  unsigned short* himg = wimg + (size_t)NT * 2 * 128 * KP;
  float* hstate = reinterpret_cast<float*>(himg + (size_t)2 * groups * 64 * KP);
  float* keyT = hstate + (size_t)KP * Bp;
  float* queryT = keyT + (size_t)KP * Bp;
  CUtensorMap mw, mh;
  if (!map16(enc, &mw, wimg, (long long)NT * 2 * 128, KP, 128) || !map16(enc, &mh, himg, (long long)2 * groups * 64, KP, 64))
    return -1;
  const size_t smem = (size_t)ST_NSTG * ST_STAGE + ((size_t)3 * ST_U * ST_SEQ + (size_t)3 * ST_U * a.W + 4 * ST_U +
                                                    (size_t)ST_SEQ * a.W + 2) * sizeof(float) +
                      (2 * ST_NSTG + 1) * 8 + 16 + 1024;
  if (smem > 227 * 1024) return -1;
  SG_CUDA(cudaFuncSetAttribute(gru_step_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (!reuse_img) {
    const long long total = (long long)NT * 128 * KP;
    step_pack_whh_kernel<<<(int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16), 256, 0, st>>>(
        a.w_hh, wimg, N, NT, KP);
    SG_LAUNCH_CHECK("step_pack_whh_kernel");
  }
  SG_CUDA(cudaMemsetAsync(himg, 0, (size_t)2 * groups * 64 * KP * 2, st));      // h_{-1} = 0 and zero padding
  StepArgs t = {};
  t.a = a; t.hstate = hstate; t.keyT = keyT; t.queryT = queryT; t.himg = himg; t.NCH = NCH; t.KP = KP; t.Bp = Bp;
  ProfileHook* hook = profile_hook();
  if (hook->start != nullptr) SG_CUDA(cudaEventRecord(hook->start, st));
  dim3 grid(NT, groups);
  for (int s = 0; s < N; ++s) {
    t.s = s;
    gru_step_tc_kernel<<<grid, ST_THREADS, smem, st>>>(mw, mh, t);
    count_launch();
  }
  SG_LAUNCH_CHECK("gru_step_tc_kernel");
  if (hook->stop != nullptr) SG_CUDA(cudaEventRecord(hook->stop, st));
  step_finish_kernel<<<ceil_div(B * N, 256), 256, 0, st>>>(keyT, queryT, a.key, a.query, B, N, Bp);
  SG_LAUNCH_CHECK("step_finish_kernel");
  return 0;
}

}  // namespace sg
