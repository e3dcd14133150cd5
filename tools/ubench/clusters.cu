// Micro-benchmark (debug tool, not part of libvtc): how many SMs a persistent one-CTA-per-SM kernel can cover for a given cluster size
// (clusters live inside one GPC, so a cluster size that does not divide a GPC's SM count strands SMs).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/clusters_bench tools/ubench/clusters.cu && tools/ubench/clusters_bench
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k(int* out) {
    extern __shared__ int sm[];
    if (threadIdx.x == 0) out[blockIdx.x] = sm[0];
}

int main() {
    const int smem = 200 * 1024;        // the GEMM's footprint: one CTA per SM
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    printf("%s: %d SMs\n", prop.name, prop.multiProcessorCount);
    for (int cs : {1, 2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(prop.multiProcessorCount / cs * cs);
        cfg.blockDim = dim3(256);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cs;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        int n = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
        printf("cluster size %2d: %3d co-resident clusters = %3d SMs (%s)\n", cs, n, n * cs, cudaGetErrorString(e));
    }
    return 0;
}
