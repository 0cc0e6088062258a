// fp64_pipes_probe.cu -- what bounds fp64 work on this part: the DFMA pipe, the fp64 tensor path (DMMA), or can the two
// overlap?  Decides whether the coupling-layer convolutions belong on mma.sync f64 (north star: "tensor cores only if
// ncu shows them compute-bound").  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fp64_pipes_probe ...
#include <cuda_runtime.h>
#include <stdio.h>

#define NACC 8

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double* d, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double* d, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// mode 0: DFMA only (16 chains); 1: DMMA m8n8k4 (NACC accumulator pairs); 2: both interleaved in every warp;
// 3: even warps DFMA, odd warps DMMA; 4: m16n8k8; 5: m16n8k16
template <int mode> __global__ void __launch_bounds__(512) probe(double* out, int iters) {
    const double x = 1.0 + 1e-9 * threadIdx.x, y = 1e-12 * blockIdx.x;
    double f[16], d[2 * NACC], e[8][4];
    double av[8], bv[4];
    for (int i = 0; i < 8; ++i) av[i] = x + i;
    for (int i = 0; i < 4; ++i) bv[i] = y + i;
#pragma unroll
    for (int i = 0; i < 16; ++i) f[i] = 1.0 + i;
#pragma unroll
    for (int i = 0; i < 2 * NACC; ++i) d[i] = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) e[i][j] = 0.0;
    const bool do_fma = mode == 0 || mode == 2 || (mode == 3 && ((threadIdx.x >> 5) & 1) == 0);
    const bool do_mma = mode == 1 || mode == 2 || (mode == 3 && ((threadIdx.x >> 5) & 1) == 1);
    if (mode <= 3) {
        for (int it = 0; it < iters; ++it) {
            if (do_fma) {
#pragma unroll
                for (int i = 0; i < 16; ++i) f[i] = fma(f[i], x, y);
            }
            if (do_mma) {
#pragma unroll
                for (int i = 0; i < NACC; ++i) dmma884(d[2 * i], d[2 * i + 1], x, y);
            }
        }
    } else if (mode == 4) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma1688(e[i], av, bv);
        }
    } else {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma16816(e[i], av, bv);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += f[i];
#pragma unroll
    for (int i = 0; i < 2 * NACC; ++i) s += d[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += e[i][j];
    if (s == 123.456) out[0] = s;
}

static void launch(int mode, int blocks, int threads, double* out, int iters) {
    switch (mode) {
    case 0: probe<0><<<blocks, threads>>>(out, iters); break;
    case 1: probe<1><<<blocks, threads>>>(out, iters); break;
    case 2: probe<2><<<blocks, threads>>>(out, iters); break;
    case 3: probe<3><<<blocks, threads>>>(out, iters); break;
    case 4: probe<4><<<blocks, threads>>>(out, iters); break;
    default: probe<5><<<blocks, threads>>>(out, iters); break;
    }
}

int main() {
    double* out; cudaMalloc(&out, 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    const char* names[6] = { "DFMA only", "DMMA m8n8k4 only", "DFMA+DMMA same warp", "DFMA / DMMA alternate warps", "DMMA m16n8k8", "DMMA m16n8k16" };
    for (int threads = 128; threads <= 512; threads *= 2)
        for (int mode = 0; mode < 6; ++mode) {
            const int blocks = 148;
            launch(mode, blocks, threads, out, 100);
            cudaDeviceSynchronize();
            float best = 1e30f;
            for (int r = 0; r < 3; ++r) {
                cudaEventRecord(e0);
                launch(mode, blocks, threads, out, iters);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            const double warps = (double)blocks * threads / 32.0;
            double fma_flop = 2.0 * 16 * 32 * iters, mma_flop = 2.0 * 8 * 8 * 4 * NACC * iters;   // per warp
            double flop = 0;
            if (mode == 0) flop = warps * fma_flop;
            else if (mode == 1) flop = warps * mma_flop;
            else if (mode == 2) flop = warps * (fma_flop + mma_flop);
            else if (mode == 3) flop = warps / 2 * (fma_flop + mma_flop);
            else if (mode == 4) flop = warps * 2.0 * 16 * 8 * 8 * 8 * iters;
            else flop = warps * 2.0 * 16 * 8 * 16 * 8 * iters;
            cudaError_t err = cudaGetLastError();
            printf("threads/CTA %3d  %-30s %8.3f ms  %7.2f TFLOP/s  %s\n", threads, names[mode], best, flop / (best * 1e-3) / 1e12,
                   err == cudaSuccess ? "" : cudaGetErrorString(err));
        }
    return 0;
}
