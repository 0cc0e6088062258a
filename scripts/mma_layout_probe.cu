// verifies the fragment layout of mma.sync.m16n8k8 f64 on the device: D(16x8) = A(16x8) * B(8x8) + C   (profiles/r2_mma_layout_probe.txt)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o mma_layout_probe scripts/mma_layout_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(const double* A, const double* B, double* D) {
    const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    double a0 = A[g * 8 + t], a1 = A[(g + 8) * 8 + t], a2 = A[g * 8 + t + 4], a3 = A[(g + 8) * 8 + t + 4];
    double b0 = B[t * 8 + g], b1 = B[(t + 4) * 8 + g];          // B[k][n]
    double c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c0), "+d"(c1), "+d"(c2), "+d"(c3) : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
    D[g * 8 + 2 * t] = c0; D[g * 8 + 2 * t + 1] = c1; D[(g + 8) * 8 + 2 * t] = c2; D[(g + 8) * 8 + 2 * t + 1] = c3;
}
int main() {
    double hA[128], hB[64], hD[128], ref[128];
    for (int i = 0; i < 128; ++i) hA[i] = (i * 37 % 23) - 11 + 0.25 * (i % 5);
    for (int i = 0; i < 64; ++i) hB[i] = (i * 17 % 13) - 6 + 0.5 * (i % 3);
    for (int i = 0; i < 16; ++i) for (int n = 0; n < 8; ++n) { double s = 0; for (int kk = 0; kk < 8; ++kk) s += hA[i * 8 + kk] * hB[kk * 8 + n]; ref[i * 8 + n] = s; }
    double *dA, *dB, *dD; cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dD, sizeof hD);
    cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
    k<<<1, 32>>>(dA, dB, dD); cudaMemcpy(hD, dD, sizeof hD, cudaMemcpyDeviceToHost);
    double err = 0; for (int i = 0; i < 128; ++i) err = fmax(err, fabs(hD[i] - ref[i]));
    printf("m16n8k8 f64 layout check: max |D - ref| = %g (%s)\n", err, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
