// Microbenchmark (measurement tool, not product): latency / throughput of small tcgen05.mma (kind::f16, M=128)
// in SS mode (A from shared memory) and TS mode (A from tensor memory), dependent accumulation chains vs independent
// accumulators, N = 16/32/64.  One CTA; thread 0 issues `n` MMAs, commits to an mbarrier and waits; clock64 around it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/bin/mma_probe tools/probe/mma_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t a) {
  uint64_t d = 0;
  d |= (uint64_t)((a >> 4) & 0x3FFFu);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); }
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t id, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(s_u32(bar)), "r"(parity) : "memory");
}

// mode: 0 SS, 1 TS;  nacc: number of independent accumulators (round robin);  n: MMAs;  ncol: N;  fence_every: 0 = never
__global__ void __launch_bounds__(160, 1) probe(long long* out, int mode, int nacc, int n, int ncol, int fence_every,
                                                int a_distinct) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* A = smem;                 // 8 chunks x 16 KB
  uint8_t* Bt = smem + 8 * 16384;    // 8 KB
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (8 * 16384 + 8192) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  if (threadIdx.x >= 32) {           // zero the TMEM A region (columns 256..511)
    const int w = (threadIdx.x >> 5) & 3;
    for (int c = 256; c < 512; c += 8)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};"
                   ::"r"(tm + ((uint32_t)(w * 32) << 16) + c), "r"(0u) : "memory");
    for (int c = 0; c < 256; c += 8)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};"
                   ::"r"(tm + ((uint32_t)(w * 32) << 16) + c), "r"(0u) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x == 0) {
    const uint32_t id = idesc(ncol);
    uint32_t par = 0;
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      for (int i = 0; i < n; ++i) {
        const int acc = i % nacc;
        const uint32_t d = tm + (uint32_t)acc * (uint32_t)ncol;          // accumulators in columns [0, 256)
        const uint64_t bd = desc_sw128(s_u32(Bt) + (uint32_t)(i & 3) * 32u);
        const uint32_t accum = i >= nacc ? 1u : 0u;
        if (mode == 0) {
          const int kk = a_distinct ? i : 0;
          mma_ss(d, desc_sw128(s_u32(A) + (uint32_t)((kk >> 2) & 7) * 16384u + (uint32_t)(kk & 3) * 32u), bd, id, accum);
        } else {
          const int kk = a_distinct ? (i & 31) : 0;
          mma_ts(d, tm + 256u + (uint32_t)kk * 8u, bd, id, accum);
        }
        if (fence_every > 0 && (i + 1) % fence_every == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
      }
      const long long t1 = clock64();
      commit(&bar);
      mbar_wait(&bar, par);
      par ^= 1u;
      const long long t2 = clock64();
      out[rep * 2 + 0] = t1 - t0;
      out[rep * 2 + 1] = t2 - t0;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
  }
}

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred;
}

// warp-converged issue: the whole warp 0 runs the loop, one elected lane issues.  UNR = compile-time unroll (descriptor
// offsets become constants).  TS mode, N=16, 2 accumulators.
template <int UNR, int MODE>
__global__ void __launch_bounds__(160, 1) probe2(long long* out, int n, int ncol) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* A = smem;
  uint8_t* Bt = smem + 8 * 16384;
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (8 * 16384 + 8192) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  if (threadIdx.x >= 32) {
    const int w = (threadIdx.x >> 5) & 3;
    for (int c = 0; c < 512; c += 8)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};"
                   ::"r"(tm + ((uint32_t)(w * 32) << 16) + c), "r"(0u) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x < 32) {
    const uint32_t id = idesc(ncol);
    const uint32_t b_base = s_u32(Bt), a_base = s_u32(A);
    uint32_t par = 0;
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      for (int i0 = 0; i0 < n; i0 += UNR) {
        if (elect_one()) {
#pragma unroll
          for (int u = 0; u < UNR; ++u) {
            const uint64_t bd = desc_sw128(b_base + (uint32_t)(u & 3) * 32u);
            const uint32_t d = tm + (uint32_t)(u & 1) * 16u;
            const uint32_t accum = (i0 + u) >= 2 ? 1u : 0u;
            if (MODE == 1) mma_ts(d, tm + 256u + (uint32_t)(u & 15) * 8u, bd, id, accum);
            else mma_ss(d, desc_sw128(a_base + (uint32_t)((u >> 2) & 7) * 16384u + (uint32_t)(u & 3) * 32u), bd, id, accum);
          }
        }
        __syncwarp();
      }
      const long long t1 = clock64();
      if (elect_one()) commit(&bar);
      __syncwarp();
      mbar_wait(&bar, par);
      par ^= 1u;
      const long long t2 = clock64();
      if (threadIdx.x == 0) { out[rep * 2 + 0] = t1 - t0; out[rep * 2 + 1] = t2 - t0; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
  }
}

// probe3: the GRU kernel's exact MMA pattern (23 K steps x {W_hi, W_lo}, TS mode, N=16, D at cols 0/16) with knobs:
//   a_col0 = first TMEM column of A;  data = 0 zeros / 1 normal fp16 values / 2 values incl. fp16 subnormals;
//   spin = 1: the other four warps poll an mbarrier while the MMAs run (as the epilogue warps do in the kernel)
__global__ void __launch_bounds__(160, 1) probe3(long long* out, int a_col0, int data, int spin, int nks) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* Bt = smem;                // 6 chunks x 2 KB
  __shared__ uint64_t bar, spinbar;
  __shared__ uint32_t slot;
  uint32_t seed = 12345u + threadIdx.x * 7919u;
  auto rnd16 = [&]() -> uint32_t {   // one fp16 bit pattern
    seed = seed * 1664525u + 1013904223u;
    const uint32_t r = seed >> 16;
    if (data == 0) return 0u;
    if (data == 1) return (r & 0x83FFu) | 0x3000u;            // normal values around 0.1 .. 0.25
    return ((r & 0x3u) == 0) ? (r & 0x83FFu) : ((r & 0x83FFu) | 0x3000u);   // 25 % subnormals
  };
  for (int i = threadIdx.x; i < 6 * 2048 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = rnd16() | (rnd16() << 16);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s_u32(&spinbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  if (threadIdx.x >= 32) {
    const int w = (threadIdx.x >> 5) & 3;
    for (int c = 0; c < 512; c += 8) {
      uint32_t v[8];
      for (int k = 0; k < 8; ++k) v[k] = rnd16() | (rnd16() << 16);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                   ::"r"(tm + ((uint32_t)(w * 32) << 16) + c), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]),
                   "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (threadIdx.x < 32) {
    const uint32_t id = idesc(16);
    const uint32_t b_base = s_u32(Bt);
    uint32_t par = 0;
    for (int rep = 0; rep < 3; ++rep) {
      const long long t0 = clock64();
      if (elect_one()) {
        for (int j = 0; j < nks; ++j) {
          const uint64_t bd = desc_sw128(b_base + (uint32_t)(j >> 2) * 2048u + (uint32_t)(j & 3) * 32u);
          mma_ts(tm, tm + (uint32_t)a_col0 + (uint32_t)j * 8u, bd, id, j > 0 ? 1u : 0u);
          mma_ts(tm + 16, tm + (uint32_t)a_col0 + (uint32_t)(nks + j) * 8u, bd, id, j > 0 ? 1u : 0u);
        }
        commit(&bar);
      }
      __syncwarp();
      const long long t1 = clock64();
      mbar_wait(&bar, par);
      par ^= 1u;
      const long long t2 = clock64();
      if (threadIdx.x == 0) { out[rep * 2 + 0] = t1 - t0; out[rep * 2 + 1] = t2 - t0; }
    }
    if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(&spinbar)) : "memory");
  } else if (spin) {
    mbar_wait(&spinbar, 0);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
  }
}

void run3(long long* d, size_t smem, int a_col0, int data, int spin, int nks) {
  cudaFuncSetAttribute(probe3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe3<<<1, 160, smem>>>(d, a_col0, data, spin, nks);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[6];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("gru-pattern TS a_col0=%3d data=%d spin=%d nks=%d : issue %lld  total %lld cycles  (%.1f cyc/MMA)  [%s]\n", a_col0, data,
         spin, nks, h[4], h[5], (double)h[5] / (2 * nks), cudaGetErrorString(e));
}

template <int UNR, int MODE>
void run2(long long* d, size_t smem, int n, int ncol) {
  cudaFuncSetAttribute(probe2<UNR, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe2<UNR, MODE><<<1, 160, smem>>>(d, n, ncol);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[6];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("converged %s UNR=%2d n=%2d N=%3d : issue %lld  total %lld cycles  (%.1f cyc/MMA)  [%s]\n", MODE ? "TS" : "SS", UNR, n,
         ncol, h[4], h[5], (double)h[5] / n, cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  const size_t smem = 8 * 16384 + 8192 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  struct Cfg { int mode, nacc, n, ncol, fence_every, a_distinct; };
  const Cfg cfgs[] = {
      {0, 1, 8, 16, 0, 1},  {0, 1, 32, 16, 0, 1}, {0, 2, 32, 16, 0, 1}, {0, 4, 32, 16, 0, 1}, {0, 8, 32, 16, 0, 1},
      {0, 8, 64, 16, 0, 1}, {0, 8, 64, 16, 0, 0},
      {1, 1, 8, 16, 0, 1},  {1, 1, 32, 16, 0, 1}, {1, 2, 32, 16, 0, 1}, {1, 4, 32, 16, 0, 1}, {1, 8, 32, 16, 0, 1},
      {1, 8, 64, 16, 0, 1}, {1, 8, 64, 16, 0, 0}, {1, 1, 32, 16, 0, 0},
      {1, 1, 32, 32, 0, 1}, {1, 1, 32, 64, 0, 1},
      {1, 8, 32, 16, 4, 1}, {0, 8, 32, 16, 4, 1}, {1, 1, 32, 16, 1, 1},
  };
  run3(d, smem, 32, 0, 0, 23);
  run3(d, smem, 32, 1, 0, 23);
  run3(d, smem, 32, 2, 0, 23);
  run3(d, smem, 32, 1, 1, 23);
  run3(d, smem, 128, 1, 0, 23);
  run3(d, smem, 128, 2, 1, 23);
  run3(d, smem, 32, 1, 0, 10);
  run2<1, 1>(d, smem, 32, 16);
  run2<2, 1>(d, smem, 32, 16);
  run2<8, 1>(d, smem, 32, 16);
  run2<16, 1>(d, smem, 64, 16);
  run2<16, 1>(d, smem, 64, 64);
  run2<1, 0>(d, smem, 32, 16);
  run2<8, 0>(d, smem, 32, 16);
  run2<16, 0>(d, smem, 64, 16);
  for (const Cfg& c : cfgs) {
    probe<<<1, 160, smem>>>(d, c.mode, c.nacc, c.n, c.ncol, c.fence_every, c.a_distinct);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[6];
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%s nacc=%d n=%2d N=%3d fence_every=%d a_distinct=%d : issue %lld  total %lld cycles  (%.1f cyc/MMA)  [%s]\n",
           c.mode ? "TS" : "SS", c.nacc, c.n, c.ncol, c.fence_every, c.a_distinct, h[4], h[5], (double)h[5] / c.n,
           cudaGetErrorString(e));
  }
  return 0;
}
