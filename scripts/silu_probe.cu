// silu_probe.cu -- what bounds the SiLU passes of the resident-chain kernel?  One 256-thread CTA per SM (as in k_chain),
// every thread activates blocks of U of its own shared-memory values in place exactly like act_pass (exp_fast + cubic
// reciprocal + derivative), with pieces switched off one at a time.  Prints fp64-pipe cycles per element per warp
// (pipe time alone: 19 fp64 operations x 2 cycles = 38).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false --expt-relaxed-constexpr -I fthmc_b200/csrc -o build/silu_probe scripts/silu_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include "chain_engine.cuh"
using namespace fthmc;

// FLAGS: 1 = table lookup, 2 = hardware reciprocal seed, 4 = global store of the derivative, 8 = shared-memory load/store
template <int FLAGS, int U>
__global__ void __launch_bounds__(256, 1) k(double* gout, int iters, int n_per_thread) {
    extern __shared__ __align__(16) double fthmc_dyn_smem[];
    for (int i = threadIdx.x; i < 64; i += blockDim.x) fthmc_dyn_smem[FT_EXP_TAB_OFF + i] = c_exp_tab[i];
    double* buf = fthmc_dyn_smem + FT_SMEM_PREFIX;
    for (int i = threadIdx.x; i < n_per_thread * 256; i += 256) buf[i] = -3.0 + 6.0 * ((i * 37) % 1024) / 1024.0;
    __syncthreads();
    double* dsave = gout + (size_t)blockIdx.x * n_per_thread * 256;
    double carry = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
        for (int e0 = 0; e0 < n_per_thread; e0 += U) {
            int idx[U]; double z[U], h[U], d[U];
#pragma unroll
            for (int j = 0; j < U; ++j) {
                idx[j] = (e0 + j) * 256 + threadIdx.x;
                z[j] = (FLAGS & 8) ? buf[idx[j]] : carry + 1e-3 * (e0 + j);
            }
#pragma unroll
            for (int j = 0; j < U; ++j) {
                // exp_fast(-z) with the optional pieces
                const double x = -z[j];
                const double* K = c_exp;
                const double nm = fma(x, K[4], K[5]);
                const double n = nm - K[5];
                double r = fma(n, K[6], x);
                r = fma(n, K[7], r);
                double w = K[0];
#pragma unroll
                for (int i = 1; i < 4; ++i) w = fma(w, r, K[i]);
                const double q = r * fma(r, w, 1.0);
                const int ni = __double2loint(nm);
                double tj = (FLAGS & 1) ? fthmc_dyn_smem[FT_EXP_TAB_OFF + (ni & 63)] : 1.0 + 1e-9 * (ni & 63);
                int kk = ni >> 6;
                kk = kk < -1022 ? -1022 : (kk > 1021 ? 1021 : kk);
                const double sc = __hiloint2double(__double2hiint(tj) + (kk << 20), __double2loint(tj));
                const double ex = fma(sc, q, sc);
                const double den = 1.0 + ex;
                double y;
                if (FLAGS & 2) asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(den));
                else y = 2.0 - den * 0.25;                  // (same operation count class: one DFMA instead of the MUFU)
                const double e = fma(-den, y, 1.0);
                y = fma(y, fma(e, e, e), y);
                h[j] = z[j] * y;
                d[j] = fma(y, fma(-z[j], y, z[j]), y);
            }
#pragma unroll
            for (int j = 0; j < U; ++j) {
                if (FLAGS & 8) buf[idx[j]] = h[j]; else carry += h[j] * 1e-30;
                if (FLAGS & 4) dsave[idx[j]] = d[j]; else carry += d[j] * 1e-30;
            }
        }
    }
    if (carry == 123.456) gout[0] = carry;
}

template <int FLAGS, int U> void run(const char* name, double* gout, int smem) {
    const int iters = 200, npt = 32;
    cudaFuncSetAttribute(k<FLAGS, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<FLAGS, U><<<148, 256, smem>>>(gout, 2, npt); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); k<FLAGS, U><<<148, 256, smem>>>(gout, iters, npt); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    // per SM sub-partition: 2 warps x npt elements per iteration
    const double cyc = best * 1e-3 * 1.965e9 / ((double)iters * npt * 2);
    printf("%-58s U=%d  %6.1f cycles per warp-element per scheduler  (%s)\n", name, U, cyc, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    double* gout; cudaMalloc(&gout, (size_t)148 * 32 * 256 * 8);
    const int smem = 200 * 1024;
    run<15, 4>("full: table + MUFU seed + global store + smem in/out", gout, smem);
    run<15, 8>("full", gout, smem);
    run<14, 4>("no table lookup", gout, smem);
    run<13, 4>("no MUFU seed", gout, smem);
    run<11, 4>("no global store", gout, smem);
    run<7, 4>("no smem load/store", gout, smem);
    run<0, 4>("arithmetic only", gout, smem);
    run<0, 8>("arithmetic only", gout, smem);
    run<8, 4>("smem in/out only extras", gout, smem);
    run<1, 4>("table only extras", gout, smem);
    run<2, 4>("MUFU only extras", gout, smem);
    run<4, 4>("global store only extras", gout, smem);
    return 0;
}
