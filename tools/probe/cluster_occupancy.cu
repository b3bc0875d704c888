#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256,1) kern(int* p) { extern __shared__ double s[]; if (p) p[0] = (int)s[0]; }
int main() {
    for (int cs : {2, 4, 8, 16}) {
        for (size_t smem : {(size_t)72000, (size_t)150000}) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (cs > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(cs * 32); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int n = -1;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
            printf("cluster %d smem %zu: max active clusters %d (%s)\n", cs, smem, n, cudaGetErrorString(e));
        }
    }
    return 0;
}
