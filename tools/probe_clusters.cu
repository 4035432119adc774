// How many thread-block clusters of 2/4/6/8 CTAs with ~220 KB of dynamic shared memory can be co-resident on this device?
// nvcc -arch=sm_100a -o /tmp/probe_clusters tools/probe_clusters.cu && /tmp/probe_clusters
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { if (p) *p = 1; }
int main() {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 3, 4, 6, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 64, 1, 1); cfg.blockDim = dim3(384, 1, 1); cfg.dynamicSmemBytes = 220 * 1024;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("cluster size %2d: max active clusters %3d (%3d SMs)  %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
