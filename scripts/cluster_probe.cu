// cluster_probe.cu -- can this part co-schedule clusters of 8 / 16 CTAs that each take a whole SM's shared memory
// (the L=64..128 resident-chain path), how many at once, and what does a cluster barrier / a DSMEM load cost?
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdio.h>
namespace cg = cooperative_groups;

__global__ void __launch_bounds__(256, 1) k(double* out, int iters, int mode) {
    extern __shared__ __align__(16) double sm[];
    cg::cluster_group cl = cg::this_cluster();
    const unsigned r = cl.block_rank(), n = cl.num_blocks();
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = r * 1000.0 + i;
    cl.sync();
    double acc = 0.0;
    long long t0 = clock64();
    if (mode == 0) {                 // cluster barriers
        for (int it = 0; it < iters; ++it) { cl.sync(); }
    } else if (mode == 1) {          // dependent DSMEM loads from the next rank
        const double* peer = cl.map_shared_rank(sm, (r + 1) % n);
        int idx = threadIdx.x;
        for (int it = 0; it < iters; ++it) { double v = peer[idx & 1023]; acc += v; idx = (int)v & 1023; }
    } else if (mode == 2) {          // streaming DSMEM loads (8 independent per iteration)
        const double* peer = cl.map_shared_rank(sm, (r + 1) % n);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += peer[(threadIdx.x + 32 * j + it) & 1023];
        }
    } else {                         // local __syncthreads
        for (int it = 0; it < iters; ++it) { __syncthreads(); }
    }
    long long t1 = clock64();
    cl.sync();
    if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = (double)(t1 - t0) / iters; out[1] = acc; }
}

int main() {
    double* out; cudaMalloc(&out, 64);
    const size_t smem = 226 * 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int nc = 1; nc <= 16; nc *= 2) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(nc * 4); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = nc; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int maxc = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&maxc, k, &cfg);
        printf("cluster %2d x 226 KiB: max active clusters %d (%s)\n", nc, maxc, cudaGetErrorString(e));
        const char* names[4] = { "cluster.sync", "dependent DSMEM load", "DSMEM load x8 ILP (per iteration)", "__syncthreads" };
        for (int mode = 0; mode < 4; ++mode) {
            int iters = 2000;
            e = cudaLaunchKernelEx(&cfg, k, out, iters, mode);
            cudaError_t e2 = cudaDeviceSynchronize();
            double h[2] = { 0, 0 };
            cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
            printf("   %-36s %8.1f cycles  (%s / %s)\n", names[mode], h[0], cudaGetErrorString(e), cudaGetErrorString(e2));
        }
    }
    return 0;
}
