// const_weight_probe.cu -- are the small convolutions of the resident-chain kernel faster with their warp-uniform weights
// read from CONSTANT memory (LDC through the constant cache) instead of shared memory (LDS broadcast: the phases are bound
// by shared-memory wavefronts)?  Mimics conv3 of one L = 32 chain (8 stripe groups x 32 rows of active sites, 8 input
// channels on 3 columns, 3 outputs) for 24 layers per iteration, one 256-thread CTA per SM as in k_chain.
//   variant 0: weights in shared memory, a PAIR of sites per thread on 4 warps   (the shipped form)
//   variant 1: weights in constant memory (layer-dependent base), pair per thread on 4 warps
//   variant 2: weights in constant memory, ONE site per thread on all 8 warps
//   variant 3: weights in shared memory, one site per thread on all 8 warps
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -o build/const_weight_probe scripts/const_weight_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>

constexpr int NH = 8, R = 32, G = 8, V = 1024, PAD = 2, sB = 3 * V / 4 + PAD, T = G * R, NL = 24, WSTRIDE = 320;
__constant__ double cw[NL * WSTRIDE];

struct alignas(16) dbl2 { double x, y; };
__device__ __forceinline__ dbl2 ld2(const double* p) { return *reinterpret_cast<const dbl2*>(p); }
__device__ __forceinline__ void st2(double* p, double x, double y) { dbl2 v; v.x = x; v.y = y; *reinterpret_cast<dbl2*>(p) = v; }

template <bool CONST>
__device__ __forceinline__ void conv3_pair(const double* B, const double* W, double* OUT, int gi, int r) {
    const int rm = r == 0 ? R - 1 : r - 1, rp = r + 2 == R ? 0 : r + 2;
    double o0[3][3], o1[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { o0[a][0] = o0[a][1] = o0[a][2] = 0.0; o1[a][0] = o1[a][1] = o1[a][2] = 0.0; }
#pragma unroll 2
    for (int ci = 0; ci < NH; ++ci) {
        const double* Bp = B + ci * sB + 3 * gi * R;
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const dbl2 v12 = ld2(Bp + b * R + r);
            const double v[4] = { Bp[b * R + rm], v12.x, v12.y, Bp[b * R + rp] };
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                double w0, w1, w2;
                if (CONST) { w0 = W[((ci * 3 + a) * 3 + b) * 4]; w1 = W[((ci * 3 + a) * 3 + b) * 4 + 1]; w2 = W[((ci * 3 + a) * 3 + b) * 4 + 2]; }
                else { const dbl2 w01 = ld2(W + ((ci * 3 + a) * 3 + b) * 4); w0 = w01.x; w1 = w01.y; w2 = W[((ci * 3 + a) * 3 + b) * 4 + 2]; }
                o0[a][0] = fma(w0, v[a], o0[a][0]); o0[a][1] = fma(w1, v[a], o0[a][1]); o0[a][2] = fma(w2, v[a], o0[a][2]);
                o1[a][0] = fma(w0, v[a + 1], o1[a][0]); o1[a][1] = fma(w1, v[a + 1], o1[a][1]); o1[a][2] = fma(w2, v[a + 1], o1[a][2]);
            }
        }
    }
    const int t = gi * R + r;
#pragma unroll
    for (int o = 0; o < 3; ++o)
        st2(OUT + o * T + t, ((W[288 + o] + o0[0][o]) + o0[1][o]) + o0[2][o], ((W[288 + o] + o1[0][o]) + o1[1][o]) + o1[2][o]);
}

template <bool CONST>
__device__ __forceinline__ void conv3_one(const double* B, const double* W, double* OUT, int gi, int r) {
    const int rm = r == 0 ? R - 1 : r - 1, rp = r + 1 == R ? 0 : r + 1;
    double o0[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a) o0[a][0] = o0[a][1] = o0[a][2] = 0.0;
#pragma unroll 2
    for (int ci = 0; ci < NH; ++ci) {
        const double* Bp = B + ci * sB + 3 * gi * R;
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const double v[3] = { Bp[b * R + rm], Bp[b * R + r], Bp[b * R + rp] };
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                double w0, w1, w2;
                if (CONST) { w0 = W[((ci * 3 + a) * 3 + b) * 4]; w1 = W[((ci * 3 + a) * 3 + b) * 4 + 1]; w2 = W[((ci * 3 + a) * 3 + b) * 4 + 2]; }
                else { const dbl2 w01 = ld2(W + ((ci * 3 + a) * 3 + b) * 4); w0 = w01.x; w1 = w01.y; w2 = W[((ci * 3 + a) * 3 + b) * 4 + 2]; }
                o0[a][0] = fma(w0, v[a], o0[a][0]); o0[a][1] = fma(w1, v[a], o0[a][1]); o0[a][2] = fma(w2, v[a], o0[a][2]);
            }
        }
    }
    const int t = gi * R + r;
#pragma unroll
    for (int o = 0; o < 3; ++o) OUT[o * T + t] = ((W[288 + o] + o0[0][o]) + o0[1][o]) + o0[2][o];
}

template <int VAR>
__global__ void __launch_bounds__(256, 1) k(double* gout, long long* cyc, int iters) {
    extern __shared__ __align__(16) double sm[];
    double* B = sm; double* Ws = B + NH * sB; double* OUT = Ws + WSTRIDE;
    for (int i = threadIdx.x; i < NH * sB; i += 256) B[i] = 1e-3 * ((i * 37) % 1024);
    for (int i = threadIdx.x; i < WSTRIDE; i += 256) Ws[i] = cw[i];
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it)
#pragma unroll 1
        for (int l = 0; l < NL; ++l) {
            const double* W = (VAR == 1 || VAR == 2) ? cw + l * WSTRIDE : Ws;
            if (VAR == 0 || VAR == 1) {
                for (int t2 = threadIdx.x; t2 < T / 2; t2 += 256) { const int gi = t2 / (R / 2); conv3_pair<VAR == 1>(B, W, OUT, gi, 2 * (t2 - gi * (R / 2))); }
            } else {
                for (int t = threadIdx.x; t < T; t += 256) { const int gi = t / R; conv3_one<VAR == 2>(B, W, OUT, gi, t - gi * R); }
            }
            __syncthreads();
            if (threadIdx.x < 3) B[threadIdx.x * 5 + l] += OUT[threadIdx.x * T + l] * 1e-9;   // (a dependence between layers)
            __syncthreads();
        }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (threadIdx.x < 3) gout[blockIdx.x * 3 + threadIdx.x] = OUT[threadIdx.x * T + 7];
}

int main() {
    double h[NL * WSTRIDE];
    for (int i = 0; i < NL * WSTRIDE; ++i) h[i] = 1e-2 * ((i * 13) % 97 - 48);
    cudaMemcpyToSymbol(cw, h, sizeof(h));
    double* gout; long long* cyc;
    cudaMalloc(&gout, 148 * 3 * sizeof(double)); cudaMalloc(&cyc, 148 * sizeof(long long));
    const size_t smem = (NH * sB + WSTRIDE + 3 * T) * sizeof(double);
    const int iters = 10;
    const char* names[4] = { "smem weights, pair/thread, 4 warps (shipped)", "const weights, pair/thread, 4 warps", "const weights, site/thread, 8 warps", "smem weights, site/thread, 8 warps" };
    for (int v = 0; v < 4; ++v) {
        for (int rep = 0; rep < 2; ++rep) {
            if (v == 0) { cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<0><<<148, 256, smem>>>(gout, cyc, iters); }
            if (v == 1) { cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<1><<<148, 256, smem>>>(gout, cyc, iters); }
            if (v == 2) { cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<2><<<148, 256, smem>>>(gout, cyc, iters); }
            if (v == 3) { cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); k<3><<<148, 256, smem>>>(gout, cyc, iters); }
            cudaDeviceSynchronize();
        }
        long long hc[148]; double ho[3];
        cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost); cudaMemcpy(ho, gout, sizeof(ho), cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; ++i) avg += hc[i];
        printf("variant %d  %-48s  %8.1f kcycles per 24 layers   (check %.6e %s)\n", v, names[v], avg / 148 / iters / 1e3, ho[0], cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
