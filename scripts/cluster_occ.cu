// Diagnostics: how many clusters of 8 CTAs (576 threads, ~215 KB dynamic shared memory each) can be co-resident on this GPU?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(576, 1) k(int* out) {
  extern __shared__ unsigned char sm[];
  if (threadIdx.x == 0) { sm[0] = 1; unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); out[blockIdx.x] = (int)smid; }
}
int main() {
  for (int smem_kb : {100, 200, 215, 225}) {
    for (int cs : {2, 4, 8, 16}) {
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024);
      if (cs > 8) cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs * 16, 1, 1); cfg.blockDim = dim3(576, 1, 1); cfg.dynamicSmemBytes = smem_kb * 1024;
      cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = cs; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
      cfg.attrs = a; cfg.numAttrs = 1;
      int n = -1;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
      printf("smem %d KB cluster %2d: max active clusters %d (%s) -> %d SMs\n", smem_kb, cs, n, cudaGetErrorString(e), n * cs);
      cudaGetLastError();
    }
  }
  int* d; cudaMalloc(&d, 1024 * 4); cudaMemset(d, 0xff, 1024 * 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 215 * 1024);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(128, 1, 1); cfg.blockDim = dim3(576, 1, 1); cfg.dynamicSmemBytes = 215 * 1024;
  cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = 8; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
  cfg.attrs = a; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k, d);
  cudaDeviceSynchronize();
  int h[128]; cudaMemcpy(h, d, 128 * 4, cudaMemcpyDeviceToHost);
  printf("launch: %s; smids:", cudaGetErrorString(e));
  for (int i = 0; i < 128; ++i) printf(" %d", h[i]);
  printf("\n");
  return 0;
}
