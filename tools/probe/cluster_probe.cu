// Scratch probe: how many thread-block clusters of each size a device keeps resident (1 CTA/SM, 120 KB smem).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dummy(float* p) { extern __shared__ float s[]; if (p) p[0] = s[0]; }
int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  printf("%s SMs=%d\n", pr.name, pr.multiProcessorCount);
  const int smem = 120 * 1024;
  cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs = 1; cs <= 16; ++cs) {
    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(cs * 32); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs;
    at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1; cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
    printf("cluster size %2d: max active clusters %3d (%3d SMs) %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaGetLastError();
  }
  return 0;
}
