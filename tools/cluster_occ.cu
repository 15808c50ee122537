// How many thread-block clusters of a given size are co-resident on this GPU for a 1024-thread CTA with large shared memory?
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/bin/cluster_occ tools/cluster_occ.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(1024, 1) k(double* out) { extern __shared__ double sm[]; if (out) out[0] = sm[threadIdx.x]; }
int main() {
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int kb : {100, 170, 200, 225}) {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kb * 1024);
    for (int cl : {1, 2, 4, 5, 6, 7, 8, 16}) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cl * 64); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = size_t(kb) * 1024;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = -1;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
      printf("smem %3d KB cluster %2d: max active clusters %3d (%d CTAs) %s\n", kb, cl, n, n * cl, e == cudaSuccess ? "" : cudaGetErrorString(e));
      (void)cudaGetLastError();
    }
  }
  return 0;
}
