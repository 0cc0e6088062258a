// dmma_latency_probe.cu -- DMMA.8x8x4 throughput vs independent accumulator tiles per warp at 1 / 2 / 4 warps per scheduler:
// how many accumulator tiles ph_conv2_mma / ph_conv2T_mma must keep in flight.
#include <cuda_runtime.h>
#include <stdio.h>
template <int NACC> __global__ void __launch_bounds__(512) k(double* out, int iters) {
    const double x = 1.0 + 1e-9 * threadIdx.x, y = 1e-12 * blockIdx.x;
    double d[2 * NACC];
#pragma unroll
    for (int i = 0; i < 2 * NACC; ++i) d[i] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 12 / NACC; ++u)
#pragma unroll
            for (int i = 0; i < NACC; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d[2 * i]), "+d"(d[2 * i + 1]) : "d"(x), "d"(y));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 2 * NACC; ++i) s += d[i];
    if (s == 123.456) out[0] = s;
}
template <int NACC> float run(double* out, int threads, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NACC><<<148, threads>>>(out, 10); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); k<NACC><<<148, threads>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}
int main() {
    double* out; cudaMalloc(&out, 64);
    const int iters = 5000;
    printf("DMMA.8x8x4 TFLOP/s (12 DMMA per warp per iteration); [cycles per dependent DMMA at 1 tile, 1 warp/scheduler]\n");
    printf("%-18s %8s %8s %8s %8s %8s %8s\n", "warps/scheduler", "1 tile", "2", "3", "4", "6", "12");
    for (int threads = 128; threads <= 512; threads *= 2) {
        float ms[6] = { run<1>(out, threads, iters), run<2>(out, threads, iters), run<3>(out, threads, iters), run<4>(out, threads, iters), run<6>(out, threads, iters), run<12>(out, threads, iters) };
        printf("%-18d", threads / 128);
        for (int i = 0; i < 6; ++i) printf(" %8.2f", 2.0 * 256 * 12 * iters * (threads / 32) * 148 / (ms[i] * 1e-3) / 1e12);
        printf("   [%.1f cyc]\n", ms[0] * 1e-3 * 1.965e9 / (12.0 * iters));
    }
    return 0;
}
