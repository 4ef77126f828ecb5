// Prints how many thread-block clusters of each size can be co-resident on the device for a 1-CTA/SM kernel using
// `smem` bytes of dynamic shared memory (decides the decoder-sequence kernel's cluster layout).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* p) { extern __shared__ float s[]; if (p) p[0] = s[0]; }
int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("SMs %d\n", sms);
    for (int smem : {64 << 10, 200 << 10, 220 << 10}) {
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        for (int cs : {1, 2, 4, 8, 16}) {
            cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(288); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
            printf("smem %3d KB cluster %2d: max active clusters %d (%d CTAs) %s\n", smem >> 10, cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    }
    return 0;
}
