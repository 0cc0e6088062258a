// dfma_mac_probe.cu -- DFMA rate of a register-blocked MAC tile (acc[i][o] += in[i] * w[o], all operands in registers,
// no memory traffic) at 2 warps per scheduler, against the DMMA form of the same work: does the register file or the
// fp64 pipe bound the convolution loops?
#include <cuda_runtime.h>
#include <stdio.h>
template <int NI, int NO> __global__ void __launch_bounds__(256) kmac(double* out, int iters) {
    double in[NI], w[NO], acc[NI][NO];
#pragma unroll
    for (int i = 0; i < NI; ++i) in[i] = 1.0 + 1e-9 * (threadIdx.x + i);
#pragma unroll
    for (int o = 0; o < NO; ++o) w[o] = 1e-12 * (blockIdx.x + o);
#pragma unroll
    for (int i = 0; i < NI; ++i)
#pragma unroll
        for (int o = 0; o < NO; ++o) acc[i][o] = i + o;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NI; ++i)
#pragma unroll
            for (int o = 0; o < NO; ++o) acc[i][o] = fma(w[o], in[i], acc[i][o]);
        // rotate the operands so that the compiler cannot hoist anything
        double t = in[0];
#pragma unroll
        for (int i = 0; i + 1 < NI; ++i) in[i] = in[i + 1];
        in[NI - 1] = t;
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < NI; ++i)
#pragma unroll
        for (int o = 0; o < NO; ++o) s += acc[i][o];
    if (s == 123.456) out[0] = s;
}
template <int NI, int NO> void run(double* out, const char* name) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    kmac<NI, NO><<<148, 256>>>(out, 10); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); kmac<NI, NO><<<148, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("%-28s %7.2f TFLOP/s\n", name, 2.0 * NI * NO * iters * 256.0 * 148 / (best * 1e-3) / 1e12);
}
int main() {
    double* out; cudaMalloc(&out, 64);
    run<6, 4>(out, "tile 6 inputs x 4 weights");
    run<12, 2>(out, "tile 12 inputs x 2 weights");
    run<4, 4>(out, "tile 4 x 4");
    run<8, 2>(out, "tile 8 x 2");
    run<2, 8>(out, "tile 2 x 8");
    run<6, 8>(out, "tile 6 x 8");
    return 0;
}
