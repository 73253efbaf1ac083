// Prints cudaOccupancyMaxActiveClusters for the cluster sizes / shared-memory footprints
// the lattice fast-path kernels use (how many clusters are co-resident on this GPU).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* p) { extern __shared__ float s[]; if (p) p[0] = s[0]; }
int main() {
  int smems[] = {200 * 1024, 100 * 1024, 64 * 1024};
  int threads[] = {512, 256};
  for (int smem : smems) for (int th : threads) {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int cl : {1, 2, 3, 4, 6, 8, 16}) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cl * 64); cfg.blockDim = dim3(th); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      int n = -1;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
      printf("smem %3d KB threads %3d cluster %2d -> max active clusters %3d (CTAs %3d) %s\n",
             smem / 1024, th, cl, n, n * cl, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }
  return 0;
}
