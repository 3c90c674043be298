// How many clusters of 1, 2, 4, 8 one-CTA-per-SM CTAs (384 threads, ~200 KB of shared memory: the GEMM's shape) can be
// co-resident on this GPU: decides whether a 2 x 2 multicast cluster can use every SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o cluster_occ cluster_occ.cu && ./cluster_occ
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(384, 1) dummy(int* p) {
  extern __shared__ char sm[];
  if (p) p[0] = sm[0];
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  printf("%s: %d SMs\n", prop.name, prop.multiProcessorCount);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 64);
    cfg.blockDim = dim3(384);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
    printf("cluster size %2d: max active clusters %3d -> %3d CTAs (%s)\n", cs, n, n * cs, cudaGetErrorString(e));
  }
  return 0;
}
