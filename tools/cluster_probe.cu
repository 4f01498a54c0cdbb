// How many thread-block clusters of a given size are co-resident when every CTA needs a whole SM?
// (cudaOccupancyMaxActiveClusters; B200: 8 GPCs of unequal size, a cluster must sit inside one GPC.)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { extern __shared__ int s[]; if (p) p[0] = s[0]; }
int main() {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cl : {2, 4, 6, 8, 10, 12, 16}) {
    for (int thr : {512, 576}) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(cl * 40); cfg.blockDim = dim3(thr); cfg.dynamicSmemBytes = 220 * 1024;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = -1;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
      printf("cluster %2d x %d threads, 220 KB smem: max active clusters %d (%s) -> %d SMs\n", cl, thr, n, cudaGetErrorString(e), n * cl);
    }
  }
  return 0;
}
