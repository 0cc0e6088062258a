// dfma_ilp_probe.cu -- DFMA throughput as a function of warps per scheduler and independent chains per thread:
// how much instruction-level parallelism the resident-chain kernel (2 warps per scheduler) needs to fill the fp64 pipe.
#include <cuda_runtime.h>
#include <stdio.h>
template <int C> __global__ void __launch_bounds__(512) k(double* out, int iters) {
    const double x = 1.0 + 1e-9 * threadIdx.x, y = 1e-12 * blockIdx.x;
    double f[C];
#pragma unroll
    for (int i = 0; i < C; ++i) f[i] = 1.0 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16 / C; ++u)
#pragma unroll
            for (int i = 0; i < C; ++i) f[i] = fma(f[i], x, y);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < C; ++i) s += f[i];
    if (s == 123.456) out[0] = s;
}
template <int C> float run(double* out, int threads, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<C><<<148, threads>>>(out, 10); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); k<C><<<148, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}
int main() {
    double* out; cudaMalloc(&out, 64);
    const int iters = 20000;
    printf("DFMA TFLOP/s (148 CTAs; 16 DFMA per thread per iteration; cycles per dependent DFMA in brackets at 1 chain)\n");
    printf("%-22s %8s %8s %8s %8s %8s\n", "warps/scheduler", "1 chain", "2", "4", "8", "16");
    for (int threads = 128; threads <= 512; threads *= 2) {
        float ms[5] = { run<1>(out, threads, iters), run<2>(out, threads, iters), run<4>(out, threads, iters), run<8>(out, threads, iters), run<16>(out, threads, iters) };
        printf("%-22d", threads / 128);
        for (int i = 0; i < 5; ++i) printf(" %8.2f", 2.0 * 16 * iters * threads * 148 / (ms[i] * 1e-3) / 1e12);
        printf("   [%.1f cyc]\n", ms[0] * 1e-3 * 1.965e9 / (16.0 * iters));
    }
    return 0;
}
