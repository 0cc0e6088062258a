// fthmc_capi.cu -- sm_100a kernels and the C ABI of libfthmc_b200.so (declared in include/fthmc_b200.h).
//
//   k_chain          persistent one-CTA-per-chain kernel: every flow / trajectory / run-loop / training-gradient entry point
//   k_chain_cluster  the same engine with one thread-block cluster (up to 16 CTAs, DSMEM halos) per chain: L = 48 .. 128
//   k_grad_reduce    fixed-order sum of the per-(CTA, warp) weight-gradient slices
//   k_action_topo    streaming Wilson action / topological charge, cluster-reduced per chain  (HBM bound)
//   k_force          streaming Wilson force with a shared-memory sin(P) tile (HBM bound)
//   k_regularize     elementwise wrap
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false (FMAs are written
// explicitly in the convolutions; elementwise updates keep the reference's separate mul/add rounding).
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <new>
#include <string>
#include <vector>

#include "../../include/fthmc_b200.h"
#include "chain_programs.cuh"
#include "weight_pack.h"

using namespace fthmc;

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<unsigned long long> g_launches{0};

static int fail(int code, const char* msg) { g_err = msg; return code; }
static int cuda_fail(cudaError_t e, const char* where) {
    g_err = std::string(where) + ": " + cudaGetErrorString(e);
    return (int)e;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(e_, #call); } while (0)

extern "C" const char* fthmc_last_error_string(void) { return g_err.c_str(); }
extern "C" int fthmc_version(void) { return 100; }
extern "C" unsigned long long fthmc_launch_count(void) { return g_launches.load(); }

// ------------------------------------------------------------------------------------------------
// CTA execution policy for the chain engine
// ------------------------------------------------------------------------------------------------
#ifdef FT_PROFILE
__device__ unsigned long long g_prof[32];
#endif

#ifndef FT_TMA_STAGES
#define FT_TMA_STAGES 2
#endif
#ifndef FT_STENCIL_TMA
#define FT_STENCIL_TMA 1
#endif
#ifndef FT_FAST_TRIG
#define FT_FAST_TRIG 1
#endif
#ifndef FT_WAVE_LAUNCHES
#define FT_WAVE_LAUNCHES 1
#endif
#ifndef FT_THREADS
#define FT_THREADS 256          // threads per CTA of the resident-chain kernels
#endif

struct CtaExec {
    static constexpr bool kCluster = false;
    __host__ __device__ int rank() const { return 0; }
    __host__ __device__ int nranks() const { return 1; }
    __host__ __device__ double* peer(double* p, int) const { return p; }
#ifdef FT_PROFILE
    __device__ long long clock() const { return clock64(); }
    __device__ void prof_add(int id, long long c) const { if (threadIdx.x == 0) atomicAdd(&g_prof[id], (unsigned long long)c); }
#endif
    double* red;   // 64 doubles of shared scratch (the first 64 of the FT_SMEM_PREFIX doubles that open the dynamic shared memory)
    // base of the engine's arena.  Naming the extern __shared__ symbol here (instead of carrying a
    // generic pointer in the engine) lets nvcc emit LDS/STS rather than generic LD/ST in every phase.
    __host__ __device__ double* smem() const {
#ifdef __CUDA_ARCH__
        extern __shared__ __align__(16) double fthmc_dyn_smem[];
        return fthmc_dyn_smem + FT_SMEM_PREFIX;
#else
        return nullptr;
#endif
    }
    // transaction barriers (mbarrier) live in the last 8 doubles of the 64-double scratch block
    __host__ __device__ void bar_init(int n) const {
#ifdef __CUDA_ARCH__
        if (threadIdx.x == 0) {
            for (int i = 0; i < n; ++i)
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"((unsigned)__cvta_generic_to_shared(red + 56 + i)) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        }
        __syncthreads();
#endif
    }
    // ONE thread: TMA bulk copy global -> shared (n doubles; both addresses and the byte count multiples of 16),
    // completing on barrier `bar`
    __host__ __device__ void bulk_load(int bar, double* dst, const double* src, int n) const {
#ifdef __CUDA_ARCH__
        const unsigned mb = (unsigned)__cvta_generic_to_shared(red + 56 + bar), d = (unsigned)__cvta_generic_to_shared(dst);
        const unsigned bytes = 8u * (unsigned)n;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(mb), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                     ::"r"(d), "l"(src), "r"(bytes), "r"(mb) : "memory");
#endif
    }
    __host__ __device__ void bar_wait(int bar, int parity) const {
#ifdef __CUDA_ARCH__
        const unsigned mb = (unsigned)__cvta_generic_to_shared(red + 56 + bar);
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "WAIT_%=:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra DONE_%=;\n"
            "bra WAIT_%=;\n"
            "DONE_%=:\n"
            "}\n" ::"r"(mb), "r"((unsigned)parity) : "memory");
#endif
    }
    // make this thread's earlier generic-proxy writes (global or shared) visible to later async-proxy (bulk copy) reads
    __host__ __device__ void proxy_fence() const {
#ifdef __CUDA_ARCH__
        asm volatile("fence.proxy.async;\n" ::: "memory");
#endif
    }
    __host__ __device__ int tid() const {
#ifdef __CUDA_ARCH__
        return threadIdx.x;
#else
        return 0;
#endif
    }
    __host__ __device__ int nt() const {
#ifdef __CUDA_ARCH__
        return blockDim.x;
#else
        return 1;
#endif
    }
    // finer task split of the phases: only in builds with wide blocks (-DFT_THREADS=512), compiled out otherwise
    __host__ __device__ bool fine(int tasks) const { return FT_THREADS > 256 && nt() > tasks; }
    // warp-level fp64 tensor-core tile (DMMA.8x8x4): every thread holds its own lane's fragment elements
    static constexpr int kLanes = 1;
    __host__ __device__ bool use_mma() const { return true; }
    __host__ __device__ int warp() const {
#ifdef __CUDA_ARCH__
        return threadIdx.x >> 5;
#else
        return 0;
#endif
    }
    __host__ __device__ int nwarps() const {
#ifdef __CUDA_ARCH__
        return blockDim.x >> 5;
#else
        return 1;
#endif
    }
    __host__ __device__ int lane0() const {
#ifdef __CUDA_ARCH__
        return threadIdx.x & 31;
#else
        return 0;
#endif
    }
    __host__ __device__ void mma884(double (&d0)[1], double (&d1)[1], const double (&a)[1], const double (&b)[1]) const {
#ifdef __CUDA_ARCH__
        asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
            : "+d"(d0[0]), "+d"(d1[0]) : "d"(a[0]), "d"(b[0]));
#endif
    }
    __host__ __device__ void sync() const {
#ifdef __CUDA_ARCH__
        __syncthreads();
#endif
    }
    __host__ __device__ void lsync() const {
#ifdef __CUDA_ARCH__
        __syncthreads();
#endif
    }
    __host__ __device__ double sum(double v) const {
#ifdef __CUDA_ARCH__
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
        if ((threadIdx.x & 31) == 0) red[w] = v;
        __syncthreads();
        double t = 0.0;
        for (int i = 0; i < nw; ++i) t += red[i];
        __syncthreads();
        return t;
#else
        return v;
#endif
    }
    __host__ __device__ bool all(bool pred) const {
#ifdef __CUDA_ARCH__
        return __syncthreads_and(pred) != 0;
#else
        return pred;
#endif
    }
    __host__ __device__ double maxv(double v) const {
#ifdef __CUDA_ARCH__
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
        if ((threadIdx.x & 31) == 0) red[w] = v;
        __syncthreads();
        double t = red[0];
        for (int i = 1; i < nw; ++i) t = fmax(t, red[i]);
        __syncthreads();
        return t;
#else
        return v;
#endif
    }
};

// One thread-block cluster per chain (lattices beyond one SM's shared memory): same engine, kCluster code path.
// sync / sum / maxv are cluster-wide; peer() maps a pointer into this CTA's arena to the same offset in a peer's.
struct ClusterExec : CtaExec {
    static constexpr bool kCluster = true;
#ifdef __CUDA_ARCH__
    __device__ int rank() const { return (int)cooperative_groups::this_cluster().block_rank(); }
    __device__ int nranks() const { return (int)cooperative_groups::this_cluster().num_blocks(); }
    __device__ double* peer(double* p, int r) const { return cooperative_groups::this_cluster().map_shared_rank(p, (unsigned)r); }
    __device__ void sync() const { cooperative_groups::this_cluster().sync(); }
    // CTA stage as in CtaExec, then every rank pushes its partial into slot [32 + own rank] of every rank
    template <bool MAX> __device__ double reduce(double v) const {
        auto cl = cooperative_groups::this_cluster();
        const int nr = (int)cl.num_blocks(), rk = (int)cl.block_rank();
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const double w = __shfl_xor_sync(0xffffffffu, v, o); v = MAX ? fmax(v, w) : v + w; }
        const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
        if ((threadIdx.x & 31) == 0) red[w] = v;
        __syncthreads();
        double t = red[0];
        for (int i = 1; i < nw; ++i) t = MAX ? fmax(t, red[i]) : t + red[i];
        if ((int)threadIdx.x < nr) cl.map_shared_rank(red, threadIdx.x)[32 + rk] = t;
        cl.sync();
        double tot = red[32];
        for (int i = 1; i < nr; ++i) tot = MAX ? fmax(tot, red[32 + i]) : tot + red[32 + i];   // rank order: identical on every rank
        cl.sync();
        return tot;
    }
    __device__ double sum(double v) const { return reduce<false>(v); }
    __device__ double maxv(double v) const { return reduce<true>(v); }
    __device__ bool all(bool pred) const { return reduce<true>(pred ? 0.0 : 1.0) == 0.0; }
#else
    bool all(bool pred) const { return pred; }
    int rank() const { return 0; }
    int nranks() const { return 1; }
    double* peer(double* p, int) const { return p; }
    void sync() const {}
    double sum(double v) const { return v; }
    double maxv(double v) const { return v; }
#endif
};

// The engine object sits in STATIC shared memory, padded to a multiple of 128 bytes.  The padding is load-bearing: the
// dynamic shared memory (fthmc_dyn_smem) starts behind the static part, rounded up to the alignment of the extern array --
// and this translation unit declares extern __shared__ arrays with 16-byte (the chain kernels, the engine's helpers) and
// 128-byte alignment (the TMA stencil).  With a static size that is not a multiple of 128 the kernel body and its noinline
// device functions resolved the dynamic base differently (observed: 112 bytes apart at sizeof(Engine) = 272: the exp / atan
// tables written by the kernel prologue were read 14 slots off by the phases).  A multiple of 128 makes every view agree.
#define ENGINE_BUF_BYTES(EXEC) ((sizeof(Engine<EXEC>) + 127) / 128 * 128)
// L = 32 (BASELINE config 3) fits ONE SM with 64 bytes to spare: 232 000 B of arena + 384 B of engine object of the 232 448 B
// a block may opt into.  One more 128-byte step of the engine object would silently push L = 32 onto the 2-CTA cluster path.
static_assert(ENGINE_BUF_BYTES(CtaExec) <= 384 && ENGINE_BUF_BYTES(ClusterExec) <= 384,
              "Engine grew past 384 bytes: L=32 no longer fits a single SM's shared memory -- shrink the engine object or the arena");

__global__ void __launch_bounds__(FT_THREADS, 1) k_chain_cluster(const ChainArgs a) {
    extern __shared__ __align__(16) double fthmc_dyn_smem[];
    __shared__ __align__(128) unsigned char en_buf[ENGINE_BUF_BYTES(ClusterExec)];
    Engine<ClusterExec>* en = reinterpret_cast<Engine<ClusterExec>*>(en_buf);
    auto cl = cooperative_groups::this_cluster();
    const int nr = (int)cl.num_blocks(), cid = blockIdx.x / nr, ncl = gridDim.x / nr;
    if (threadIdx.x == 0) {
        ClusterExec ex; ex.red = fthmc_dyn_smem;
        new (en) Engine<ClusterExec>(ex, a.pr, a.ws + (size_t)cid * a.ws_stride);
        en->gW = nullptr;                                    // (the training mode runs on the single-CTA path)
    }
    for (int i = threadIdx.x; i < 64; i += blockDim.x) fthmc_dyn_smem[FT_EXP_TAB_OFF + i] = c_exp_tab[i];   // exp_fast's 2^(j/64) table (blocks may be one warp)
    if (threadIdx.x < 8) fthmc_dyn_smem[FT_ATAN_TAB_OFF + 1 + threadIdx.x] = c_atan_tab[1 + threadIdx.x];   // atan2x2_fast's theta_j
    __syncthreads();
    en->ex.bar_init(Engine<ClusterExec>::NBAR);
    if (a.pr.nlayers > 0) en->load_geom_table();
    for (int b = a.b_begin + cid; b < a.b_end; b += ncl) run_chain(*en, a, b);
    cl.sync();                                   // no CTA may exit while a peer can still address its shared memory
}

// plain HMC / leapfrog only: no flow phases in the kernel, a quarter of the registers, four CTAs per SM
__global__ void __launch_bounds__(256, 4) k_chain_plain(const ChainArgs a) {
    extern __shared__ __align__(16) double fthmc_dyn_smem[];
    __shared__ __align__(128) unsigned char en_buf[ENGINE_BUF_BYTES(CtaExec)];
    Engine<CtaExec>* en = reinterpret_cast<Engine<CtaExec>*>(en_buf);
    if (threadIdx.x == 0) {
        CtaExec ex{ fthmc_dyn_smem };
        new (en) Engine<CtaExec>(ex, a.pr, a.ws + (size_t)blockIdx.x * a.ws_stride);
        en->gW = nullptr;
    }
    __syncthreads();
    for (int b = a.b_begin + blockIdx.x; b < a.b_end; b += gridDim.x) run_chain_plain(*en, a, b);
}

__global__ void __launch_bounds__(FT_THREADS, 1) k_chain(const ChainArgs a) {
    extern __shared__ __align__(16) double fthmc_dyn_smem[];
    // the engine object lives in (static) shared memory: its members are read by every noinline phase
    __shared__ __align__(128) unsigned char en_buf[ENGINE_BUF_BYTES(CtaExec)];
    Engine<CtaExec>* en = reinterpret_cast<Engine<CtaExec>*>(en_buf);
    if (threadIdx.x == 0) {
        CtaExec ex{ fthmc_dyn_smem };
        new (en) Engine<CtaExec>(ex, a.pr, a.ws + (size_t)blockIdx.x * a.ws_stride);
        en->gW = a.gbuf ? a.gbuf + (size_t)blockIdx.x * a.gbuf_stride : nullptr;
    }
    for (int i = threadIdx.x; i < 64; i += blockDim.x) fthmc_dyn_smem[FT_EXP_TAB_OFF + i] = c_exp_tab[i];   // exp_fast's 2^(j/64) table (blocks may be one warp)
    if (threadIdx.x < 8) fthmc_dyn_smem[FT_ATAN_TAB_OFF + 1 + threadIdx.x] = c_atan_tab[1 + threadIdx.x];   // atan2x2_fast's theta_j
    __syncthreads();
    en->ex.bar_init(Engine<CtaExec>::NBAR);
    if (a.pr.nlayers > 0) en->load_geom_table();
    for (int b = a.b_begin + blockIdx.x; b < a.b_end; b += gridDim.x) run_chain(*en, a, b);
}

// ------------------------------------------------------------------------------------------------
// streaming stencils
// ------------------------------------------------------------------------------------------------
template <typename T> struct M;
template <> struct M<double> {
    static __device__ double cosv(double x) { return FT_FAST_TRIG ? fthmc::cos_fast(x) : cos(x); }
    static __device__ double cos_core(double x) { return fthmc::cos_core(x); }
    static __device__ double sin_core(double x) { return FT_FAST_SIN ? fthmc::sin_core(x) : sin(x); }
    static __device__ bool cos_in_range(double x) { return fabs(x) < 524288.0; }
    static __device__ double sinv(double x) { return fthmc::sin_force(x); }
    static __device__ double floorv(double x) { return floor(x); }
    static __device__ double fmav(double a, double b, double c) { return fma(a, b, c); }
    static __device__ double absv(double x) { return fabs(x); }
    static __device__ double modv(double x, double y) { return fmod(x, y); }
};
// cosf for the fp32 action scan: x = n pi + r by a two-term reduction (fmaf), cos x = (-1)^n cos r, one even polynomial
// (Taylor to r^12: truncation 6e-9 on |r| <= pi/2) -- 11 fp32 operations; |x| < 2^15, beyond that the library.
static __device__ __forceinline__ float cosf_pi_core(float x) {
    const float t = fmaf(x, 0.318309886f, 12582912.0f);                        // n = rint(x / pi) in the low mantissa bits
    const float n = t - 12582912.0f;
    float r = fmaf(n, -3.14159274101257324f, x);
    r = fmaf(n, 8.74227765734758577e-8f, r);
    const float s = r * r;
    float p = 1.0f / 479001600.0f;
    p = fmaf(p, s, -1.0f / 3628800.0f); p = fmaf(p, s, 1.0f / 40320.0f); p = fmaf(p, s, -1.0f / 720.0f);
    p = fmaf(p, s, 1.0f / 24.0f); p = fmaf(p, s, -0.5f); p = fmaf(p, s, 1.0f);
    return __int_as_float(__float_as_int(p) ^ (__float_as_int(t) << 31));
}
static __device__ __forceinline__ float cosf_pi(float x) { return fabsf(x) < 32768.0f ? cosf_pi_core(x) : cosf(x); }
// sinf for the fp32 force: the same reduction, sin x = (-1)^n sin r, odd Taylor polynomial to r^11 (truncation 6e-8 at pi/2)
static __device__ __forceinline__ float sinf_pi_core(float x) {
    const float t = fmaf(x, 0.318309886f, 12582912.0f);
    const float n = t - 12582912.0f;
    float r = fmaf(n, -3.14159274101257324f, x);
    r = fmaf(n, 8.74227765734758577e-8f, r);
    const float s = r * r;
    float p = -1.0f / 39916800.0f;
    p = fmaf(p, s, 1.0f / 362880.0f); p = fmaf(p, s, -1.0f / 5040.0f); p = fmaf(p, s, 1.0f / 120.0f); p = fmaf(p, s, -1.0f / 6.0f);
    p = fmaf(r * s, p, r);
    return __int_as_float(__float_as_int(p) ^ (__float_as_int(t) << 31));
}
static __device__ __forceinline__ float sinf_pi(float x) { return fabsf(x) < 32768.0f ? sinf_pi_core(x) : sinf(x); }
template <> struct M<float> {
    static __device__ float cosv(float x) { return FT_FAST_TRIG ? cosf_pi(x) : cosf(x); }
    static __device__ float cos_core(float x) { return cosf_pi_core(x); }
    static __device__ float sin_core(float x) { return sinf_pi_core(x); }
    static __device__ bool cos_in_range(float x) { return fabsf(x) < 32768.0f; }
    static __device__ float sinv(float x) { return FT_FAST_TRIG ? sinf_pi(x) : sinf(x); }
    static __device__ float floorv(float x) { return floorf(x); }
    static __device__ float fmav(float a, float b, float c) { return fmaf(a, b, c); }
    static __device__ float absv(float x) { return fabsf(x); }
    static __device__ float modv(float x, float y) { return fmodf(x, y); }
};

template <typename T>
__device__ __forceinline__ T plaq_g(const T* __restrict__ f, int L0, int L1, int n0, int n1, int order) {
    const int n0p = n0 + 1 == L0 ? 0 : n0 + 1, n1p = n1 + 1 == L1 ? 0 : n1 + 1;
    const T a = f[n0 * L1 + n1], b = f[(L0 + n0p) * L1 + n1], c = f[n0 * L1 + n1p], d = f[(L0 + n0) * L1 + n1];
    return order == 0 ? ((a + b) - c) - d : ((a - d) - c) + b;
}

// (f - pi) / 2pi without the division routine, bit for bit: with C = RN(1 / 2pi), q0 = RN(a C) is a faithful quotient, the
// remainder a - q0 * 2pi is exact in one fma, and RN(q0 + r C) is the correctly rounded quotient (Markstein).  Checked
// against the division exhaustively in fp32 (every finite float) and on 4e8 random arguments in fp64; a non-finite
// argument gives NaN either way.  The charge scan and the elementwise wrap spend ~15 instructions less per value.
template <typename T>
__device__ __forceinline__ T regularize_t(T f) {
    const T PI = (T)3.141592653589793, TP = (T)6.283185307179586, C = (T)1 / TP;
    const T a = f - PI, q0 = a * C, r = M<T>::fmav(-q0, TP, a), g = M<T>::fmav(r, C, q0);
    return TP * (g - M<T>::floorv(g) - (T)0.5);
}

// torch_wrap(x) = remainder(x+pi, 2pi) - pi
// The remainder as in the chain engine's rem_2pi: k = floor(y / 2pi) from a multiply, y - k * 2pi in ONE fma (exact whenever
// k is the true quotient: the remainder of two floating-point numbers is representable), two fix-ups for a quotient that
// rounded across an integer.  The library fmod is a bit-serial loop of ~100 instructions; the batched charge scan ran at a
// fraction of the floored one's rate with it.
template <typename T>
__device__ __forceinline__ T wrap_t(T x) {
    const T PI = (T)3.141592653589793, TP = (T)6.283185307179586, C = (T)1 / TP;
    const T y = x + PI;
    if (!(M<T>::absv(y) < (T)1e6)) {                        // far outside the range of link sums: the library path
        T r = M<T>::modv(y, TP);
        if (r != (T)0 && r < (T)0) r += TP;
        return r - PI;
    }
    const T k = M<T>::floorv(y * C);
    T r = M<T>::fmav(-k, TP, y);
    if (r < (T)0) r += TP;
    if (r >= TP) r -= TP;
    return r - PI;
}

// 16-byte vectors of link angles: the streaming stencils move VEC consecutive sites of a lattice row per thread
template <typename T> struct Vec;
template <> struct Vec<double> { static constexpr int N = 2; using type = double2; };
template <> struct Vec<float> { static constexpr int N = 4; using type = float4; };

// cos of the N plaquettes of a vector with ONE range test for all of them (the dedicated cosine's polynomial path, or -- any
// |P| beyond its reduction range -- the per-value form with the library): a branch per value put four divergence regions
// into the issue-bound fp32 scan
template <typename T, int N>
__device__ __forceinline__ void cos_vec(const T (&p)[N], T (&v)[N]) {
    bool ok = FT_FAST_TRIG;
#pragma unroll
    for (int j = 0; j < N; ++j) ok = ok && M<T>::cos_in_range(p[j]);
    if (ok) {
#pragma unroll
        for (int j = 0; j < N; ++j) v[j] = M<T>::cos_core(p[j]);
    } else {
#pragma unroll
        for (int j = 0; j < N; ++j) v[j] = M<T>::cosv(p[j]);
    }
}


// plaquettes of the N sites (n0, n1 .. n1+N-1): three 16-byte loads and one scalar instead of 4N scalar loads
template <typename T>
__device__ __forceinline__ void plaq_vec(const T* __restrict__ f, int L0, int L1, int n0, int n1, int order, T (&p)[Vec<T>::N]) {
    constexpr int N = Vec<T>::N;
    using VT = typename Vec<T>::type;
    const int n0p = n0 + 1 == L0 ? 0 : n0 + 1, n1n = n1 + N == L1 ? 0 : n1 + N;
    __align__(16) T t0[N + 1], t1[N], t1p[N];
    *reinterpret_cast<VT*>(t0) = *reinterpret_cast<const VT*>(f + (size_t)n0 * L1 + n1);
    t0[N] = f[(size_t)n0 * L1 + n1n];
    *reinterpret_cast<VT*>(t1) = *reinterpret_cast<const VT*>(f + (size_t)(L0 + n0) * L1 + n1);
    *reinterpret_cast<VT*>(t1p) = *reinterpret_cast<const VT*>(f + (size_t)(L0 + n0p) * L1 + n1);
#pragma unroll
    for (int j = 0; j < N; ++j) p[j] = order == 0 ? ((t0[j] + t1p[j]) - t0[j + 1]) - t1[j] : ((t0[j] - t1[j]) - t0[j + 1]) + t1p[j];
}

template <typename T, int N>
__device__ __forceinline__ void sin_vec(T (&p)[N]) {                      // in place, one range test per vector (as cos_vec)
    bool ok = FT_FAST_TRIG;
#pragma unroll
    for (int j = 0; j < N; ++j) ok = ok && M<T>::cos_in_range(p[j]);
    if (ok) {
#pragma unroll
        for (int j = 0; j < N; ++j) p[j] = M<T>::sin_core(p[j]);
    } else {
#pragma unroll
        for (int j = 0; j < N; ++j) p[j] = M<T>::sinv(p[j]);
    }
}

// acc += f(P_j) over the N plaquettes of a vector (f = cos / regularize / wrap by WHAT); fp32 sums the vector in fp32 first
template <typename T, int WHAT, int N>
__device__ __forceinline__ void accumulate_vec(const T (&p)[N], double& acc) {
    T v[N];
    if constexpr (WHAT == 0) cos_vec<T, N>(p, v);
    else {
#pragma unroll
        for (int j = 0; j < N; ++j) v[j] = WHAT == 1 ? regularize_t(p[j]) : wrap_t(p[j]);
    }
    if constexpr (sizeof(T) == 4) {
        T grp = (T)0;
#pragma unroll
        for (int j = 0; j < N; ++j) grp += v[j];
        acc += (double)grp;
    } else {
#pragma unroll
        for (int j = 0; j < N; ++j) acc += (double)v[j];
    }
}

// what: 0 = sum cos P (action), 1 = sum regularize(P) (floored charge), 2 = sum wrap(P) (batched charge)
// grid (nc, B), thread-block cluster (nc, 1, 1): the nc CTAs of a cluster split the rows of ONE chain, reduce in fp64,
// hand their partial sums to rank 0 through distributed shared memory, and rank 0 writes the finished per-chain value.
// One launch, no global scratch, deterministic summation order.  nc == 1 (large batches): a plain one-CTA-per-chain scan.
// PIPE (with VEC): the software-pipelined scan for FEW LARGE chains (clusters; low occupancy), see below
template <typename T, bool VEC, int WHAT, int ORDER, bool PIPE = false>
__global__ void __launch_bounds__(256) k_action_topo(const T* __restrict__ links, int L0, int L1, int rows,
                                                   double beta, int rounded, T* __restrict__ out) {
    constexpr int what = WHAT, order = ORDER;                // compile time: a run-time switch evaluates every branch's arithmetic
    __shared__ double red[8];
    __shared__ double part[16];                              // rank 0: one slot per rank of the cluster
    namespace cg = cooperative_groups;
    cg::cluster_group cl = cg::this_cluster();
    const int b = blockIdx.y, c = blockIdx.x, nc = gridDim.x;
    const T* f = links + (size_t)b * 2 * L0 * L1;
    const int r0 = c * rows, r1 = min(L0, r0 + rows);
    double acc = 0.0;
    if (VEC) {
        // one 16-byte vector of sites per thread and step; (row, column) advance without divisions
        constexpr int N = Vec<T>::N;
        const int W = L1 / N, dr = (int)blockDim.x / W, dc = (int)blockDim.x - dr * W, nr = r1 - r0;
        int row = (int)threadIdx.x / W, col = (int)threadIdx.x - row * W;
        if constexpr (PIPE) {
            // Software pipeline: the link vectors of steps k + 1 and k + 2 are loaded (raw, no arithmetic on them: an add would
            // stall the in-order warp until they land) BEFORE the cos / wrap of step k.  The compiler does not move loads across
            // the divergent slow-path branches of those functions on its own, and with one step in flight per warp a scan of a
            // few large lattices ran at HBM latency (L = 1024, 48 chains: ~2000 cycles per step and warp; pipelined +24 % / +14 %
            // for action / charge).  64 registers instead of 40: with many chains (one CTA each, full occupancy) the plain loop
            // below is the faster one (L = 256: -9 % / -15 % pipelined), so the launcher picks this form for clusters only.
            using VT = typename Vec<T>::type;
            struct Raw { VT t0, t1, t1p; T t0n; };
            auto fetch = [&](int rw, int cl, Raw& q) {
                const int n0 = r0 + rw, n1 = cl * N, n0p = n0 + 1 == L0 ? 0 : n0 + 1, n1n = n1 + N == L1 ? 0 : n1 + N;
                q.t0 = *reinterpret_cast<const VT*>(f + (size_t)n0 * L1 + n1);
                q.t0n = f[(size_t)n0 * L1 + n1n];
                q.t1 = *reinterpret_cast<const VT*>(f + (size_t)(L0 + n0) * L1 + n1);
                q.t1p = *reinterpret_cast<const VT*>(f + (size_t)(L0 + n0p) * L1 + n1);
            };
            auto step = [&]() { col += dc; row += dr; if (col >= W) { col -= W; ++row; } };
            Raw qa, qb, qc;
            bool va = row < nr, vb, vc;
            if (va) fetch(row, col, qa);
            step(); vb = row < nr;
            if (vb) fetch(row, col, qb);
            while (va) {
                step(); vc = row < nr;
                if (vc) fetch(row, col, qc);
                __align__(16) T t0[N + 1], t1[N], t1p[N];
                *reinterpret_cast<VT*>(t0) = qa.t0; t0[N] = qa.t0n;
                *reinterpret_cast<VT*>(t1) = qa.t1; *reinterpret_cast<VT*>(t1p) = qa.t1p;
                T p[N];
#pragma unroll
                for (int j = 0; j < N; ++j) p[j] = order == 0 ? ((t0[j] + t1p[j]) - t0[j + 1]) - t1[j] : ((t0[j] - t1[j]) - t0[j + 1]) + t1p[j];
                accumulate_vec<T, WHAT, N>(p, acc);
                qa = qb; qb = qc; va = vb; vb = vc;
            }
        } else {
            while (row < nr) {
                T p[N];
                plaq_vec<T>(f, L0, L1, r0 + row, col * N, order, p);
                accumulate_vec<T, WHAT, N>(p, acc);
                col += dc; row += dr;
                if (col >= W) { col -= W; ++row; }
            }
        }
    } else {
        for (int i = threadIdx.x; i < (r1 - r0) * L1; i += blockDim.x) {
            const int n0 = r0 + i / L1, n1 = i % L1;
            const T p = plaq_g(f, L0, L1, n0, n1, order);
            acc += what == 0 ? (double)M<T>::cosv(p) : (what == 1 ? (double)regularize_t(p) : (double)wrap_t(p));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) {
        for (int i = 0; i < (blockDim.x + 31) / 32; ++i) t += red[i];
        if (nc > 1) cl.map_shared_rank(part, 0)[c] = t;
    }
    if (nc > 1) {
        cl.sync();
        if (c != 0) return;
        if (threadIdx.x == 0) { t = 0.0; for (int i = 0; i < nc; ++i) t += part[i]; }
    }
    if (threadIdx.x == 0) {
        double r;
        if (what == 0) r = -beta * t;
        else if (rounded) r = floor(0.1 + t / TWO_PI_D);
        else r = t / TWO_PI_D;
        out[b] = (T)r;
    }
}

// Large batches of small lattices: ONE WARP per chain (8 chains per CTA).  No block barrier and no shared memory, and the
// unrolled scan keeps four iterations of 16-byte loads in flight per lane, where the CTA-per-chain form above is bound by
// the load -> cos -> block-reduce latency of its short-lived CTAs.  Deterministic: lane-sequential sums, then a shuffle tree.
template <typename T, int WHAT, int ORDER>
__global__ void __launch_bounds__(256) k_action_topo_warp(const T* __restrict__ links, int B, int L0, int L1,
                                                        double beta, int rounded, T* __restrict__ out) {
    constexpr int what = WHAT, order = ORDER;
    const int lane = threadIdx.x & 31, b = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (b >= B) return;
    const T* f = links + (size_t)b * 2 * L0 * L1;
    constexpr int N = Vec<T>::N;
    const int W = L1 / N, nvec = L0 * W, dr = 32 / W, dc = 32 - dr * W;
    int row = lane / W, col = lane - row * W;
    double acc = 0.0;
#pragma unroll 4
    for (int i = lane; i < nvec; i += 32) {
        T p[N];
        plaq_vec<T>(f, L0, L1, row, col * N, order, p);
        accumulate_vec<T, WHAT, N>(p, acc);
        col += dc; row += dr;
        if (col >= W) { col -= W; ++row; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        double r;
        if (what == 0) r = -beta * acc;
        else if (rounded) r = floor(0.1 + acc / TWO_PI_D);
        else r = acc / TWO_PI_D;
        out[b] = (T)r;
    }
}

// Large batches of lattices that fit a shared-memory stage (<= 32 KB per chain): persistent CTAs, each streaming its
// chains through a ring of NSTAGE shared-memory buffers with TMA bulk copies (cp.async.bulk + mbarrier), one chain per
// copy.  The copies of the next NSTAGE - 1 chains are in flight while the 256 threads reduce the current one out of
// shared memory, so the HBM stream never waits for the cos / wrap arithmetic.  The scans are issue bound after that
// (ncu: issue slots 74 % busy), so the ring is kept short (FT_TMA_STAGES = 2: six CTAs of 2 x 16 KB per SM at L = 32).  One block barrier per chain: it publishes the warps' partial sums (thread 0 adds them in warp order:
// deterministic) and frees the stage for thread 0 to refill.
#ifndef FT_TMA_MINB
#define FT_TMA_MINB 1
#endif
template <typename T, int NSTAGE, int WHAT, int ORDER>
__global__ void __launch_bounds__(256, FT_TMA_MINB) k_action_topo_tma(const T* __restrict__ links, int B, int L0, int L1, int stage_elems,
                                                       double beta, int rounded, T* __restrict__ out) {
    constexpr int what = WHAT, order = ORDER;
    extern __shared__ __align__(128) unsigned char tma_raw[];
    T* buf = reinterpret_cast<T*>(tma_raw);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(tma_raw + (size_t)NSTAGE * stage_elems * sizeof(T));
    double* red = reinterpret_cast<double*>(bars + NSTAGE);                  // [2][8] partial sums, double-buffered by chain parity
    const int tid = threadIdx.x, V = L0 * L1;
    const unsigned bytes = (unsigned)(2 * V * sizeof(T));
    const int nmine = B > (int)blockIdx.x ? (B - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    auto issue = [&](int k) {                                                // thread 0: chain k of this CTA into stage k % NSTAGE
        const int st = k % NSTAGE;
        const unsigned mb = (unsigned)__cvta_generic_to_shared(bars + st), d = (unsigned)__cvta_generic_to_shared(buf + (size_t)st * stage_elems);
        const T* src = links + (size_t)(blockIdx.x + (size_t)k * gridDim.x) * 2 * V;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(mb), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                     ::"r"(d), "l"(src), "r"(bytes), "r"(mb) : "memory");
    };
    if (tid == 0) {
        for (int i = 0; i < NSTAGE; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bars + i)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        for (int k = 0; k < NSTAGE && k < nmine; ++k) issue(k);
    }
    __syncthreads();
    constexpr int N = Vec<T>::N;
    using VT = typename Vec<T>::type;
    const int W = L1 / N, nvec = L0 * W, dr = 256 / W, dc = 256 - dr * W;
    // Every chain has the same geometry: the element offsets of this thread's (at most MAXIT: a stage holds <= 32 KB) vectors
    // are computed once, outside the chain loop -- the scans are issue bound, and the (row, column) stepping with its two
    // periodic wraps was a quarter of the loop's instructions.
    constexpr int MAXIT = 4;
    int o0[MAXIT], o0n[MAXIT], o1[MAXIT], o1p[MAXIT];
    {
        int row = tid / W, col = tid - row * W;
#pragma unroll
        for (int it = 0; it < MAXIT; ++it) {
            const int n1 = col * N, n0p = row + 1 >= L0 ? 0 : row + 1, n1n = n1 + N == L1 ? 0 : n1 + N;
            o0[it] = row * L1 + n1; o0n[it] = row * L1 + n1n; o1[it] = (L0 + row) * L1 + n1; o1p[it] = (L0 + n0p) * L1 + n1;
            col += dc; row += dr;
            if (col >= W) { col -= W; ++row; }
        }
    }
    const int nit = tid < nvec ? (nvec - tid + 255) / 256 : 0;
    for (int k = 0; k < nmine; ++k) {
        const int st = k % NSTAGE;
        {
            const unsigned mb = (unsigned)__cvta_generic_to_shared(bars + st), parity = (unsigned)((k / NSTAGE) & 1);
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "WAIT_%=:\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                "@p bra DONE_%=;\n"
                "bra WAIT_%=;\n"
                "DONE_%=:\n"
                "}\n" ::"r"(mb), "r"(parity) : "memory");
        }
        const T* f = buf + (size_t)st * stage_elems;
        double acc = 0.0;
#pragma unroll
        for (int it = 0; it < MAXIT; ++it) {
            if (it < nit) {
                __align__(16) T t0[N + 1], t1[N], t1p[N];
                *reinterpret_cast<VT*>(t0) = *reinterpret_cast<const VT*>(f + o0[it]);
                t0[N] = f[o0n[it]];
                *reinterpret_cast<VT*>(t1) = *reinterpret_cast<const VT*>(f + o1[it]);
                *reinterpret_cast<VT*>(t1p) = *reinterpret_cast<const VT*>(f + o1p[it]);
                T pp[N];
#pragma unroll
                for (int j = 0; j < N; ++j) pp[j] = order == 0 ? ((t0[j] + t1p[j]) - t0[j + 1]) - t1[j] : ((t0[j] - t1[j]) - t0[j + 1]) + t1p[j];
                accumulate_vec<T, WHAT, N>(pp, acc);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if ((tid & 31) == 0) red[(k & 1) * 8 + (tid >> 5)] = acc;
        __syncthreads();                                     // partial sums published; every thread is done with stage st
        if (tid == 0) {
            if (k + NSTAGE < nmine) issue(k + NSTAGE);
            double t = 0.0;
#pragma unroll
            for (int w = 0; w < 8; ++w) t += red[(k & 1) * 8 + w];
            double r;
            if (what == 0) r = -beta * t;
            else if (rounded) r = floor(0.1 + t / TWO_PI_D);
            else r = t / TWO_PI_D;
            out[blockIdx.x + (size_t)k * gridDim.x] = (T)r;
        }
    }
}

// grid (nchunk, B): rows [r0,r1) of one chain; sin P of rows r0-1..r1-1 staged in shared memory (one sine per site).
// VEC: 16-byte loads / stores of consecutive sites and division-free (row, column) stepping; needs L1 % Vec<T>::N == 0.
template <typename T, bool VEC>
__global__ void __launch_bounds__(256) k_force(const T* __restrict__ links, int L0, int L1, int rows, T beta, int order, T* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* S = reinterpret_cast<T*>(smem_raw);                  // (rows+1) x L1, row 0 is r0-1
    const int b = blockIdx.y, c = blockIdx.x;
    const T* f = links + (size_t)b * 2 * L0 * L1;
    T* o = out + (size_t)b * 2 * L0 * L1;
    const int r0 = c * rows, r1 = min(L0, r0 + rows), nr = r1 - r0;
    if constexpr (VEC) {
        constexpr int N = Vec<T>::N;
        using VT = typename Vec<T>::type;
        const int W = L1 / N, dr = (int)blockDim.x / W, dc = (int)blockDim.x - dr * W;
        const int row0 = (int)threadIdx.x / W, col0 = (int)threadIdx.x - row0 * W;
        for (int row = row0, col = col0; row < nr + 1;) {
            int n0 = r0 - 1 + row; if (n0 < 0) n0 += L0;
            __align__(16) T p[N];
            plaq_vec<T>(f, L0, L1, n0, col * N, order, p);
            sin_vec<T, N>(p);
            *reinterpret_cast<VT*>(S + row * L1 + col * N) = *reinterpret_cast<const VT*>(p);
            col += dc; row += dr;
            if (col >= W) { col -= W; ++row; }
        }
        __syncthreads();
        for (int row = row0, col = col0; row < nr;) {
            const int n1 = col * N;
            __align__(16) T s[N], su[N], f0[N], f1[N];
            *reinterpret_cast<VT*>(s) = *reinterpret_cast<const VT*>(S + (row + 1) * L1 + n1);
            *reinterpret_cast<VT*>(su) = *reinterpret_cast<const VT*>(S + row * L1 + n1);
            T sl = S[(row + 1) * L1 + (n1 == 0 ? L1 - 1 : n1 - 1)];
#pragma unroll
            for (int j = 0; j < N; ++j) { f0[j] = beta * (s[j] - sl); f1[j] = beta * (su[j] - s[j]); sl = s[j]; }
            *reinterpret_cast<VT*>(o + (size_t)(r0 + row) * L1 + n1) = *reinterpret_cast<const VT*>(f0);
            *reinterpret_cast<VT*>(o + (size_t)(L0 + r0 + row) * L1 + n1) = *reinterpret_cast<const VT*>(f1);
            col += dc; row += dr;
            if (col >= W) { col -= W; ++row; }
        }
    } else {
        for (int i = threadIdx.x; i < (nr + 1) * L1; i += blockDim.x) {
            int n0 = r0 - 1 + i / L1; if (n0 < 0) n0 += L0;
            S[i] = M<T>::sinv(plaq_g(f, L0, L1, n0, i % L1, order));
        }
        __syncthreads();
        for (int i = threadIdx.x; i < nr * L1; i += blockDim.x) {
            const int rr = i / L1, n1 = i % L1, n1m = n1 == 0 ? L1 - 1 : n1 - 1;
            const T s = S[(rr + 1) * L1 + n1];
            o[(r0 + rr) * L1 + n1] = beta * (s - S[(rr + 1) * L1 + n1m]);
            o[(L0 + r0 + rr) * L1 + n1] = beta * (S[rr * L1 + n1] - s);
        }
    }
}

// grid (row chunks x column chunks, B): the tile rows [r0,r1) x columns [c0,c0+cw) of one chain; sin P of rows r0-1..r1-1 and
// columns c0-1..c0+cw-1 staged in shared memory (one sine per site, plus the halo row and column).
// The column-chunked form of k_force's vector path (very wide lattices only: whole rows with the plain pitch are the faster
// layout wherever they fit, L = 32: 92 % against 84 % of HBM); needs L1 % Vec<T>::N == 0 and cw % Vec<T>::N == 0.  The staged tile has pitch cw + N: an N-wide left pad keeps the vector stores aligned, its last slot
// holds the halo column c0 - 1.  Very wide lattices are cut into column chunks (cw < L1) so that a tile of ~24 rows still
// fits ~26 KB and eight CTAs stay on an SM (whole rows of L1 = 1024 allowed 7 rows per 64 KB tile: three CTAs per SM, 62 % of HBM).
template <typename T>
__global__ void __launch_bounds__(256) k_force_tiled(const T* __restrict__ links, int L0, int L1, int rows, int cw, T beta, int order, T* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* S = reinterpret_cast<T*>(smem_raw);                  // (rows+1) x pitch, row 0 is r0-1
    const int b = blockIdx.y, ncol = L1 / cw, cr = (int)blockIdx.x / ncol, c = (int)blockIdx.x - cr * ncol;
    const T* f = links + (size_t)b * 2 * L0 * L1;
    T* o = out + (size_t)b * 2 * L0 * L1;
    const int r0 = cr * rows, r1 = min(L0, r0 + rows), nr = r1 - r0, c0 = c * cw;
    {
        constexpr int N = Vec<T>::N;
        using VT = typename Vec<T>::type;
        const int W = cw / N, SP = cw + N, dr = (int)blockDim.x / W, dc = (int)blockDim.x - dr * W;
        const int row0 = (int)threadIdx.x / W, col0 = (int)threadIdx.x - row0 * W;
        for (int row = row0, col = col0; row < nr + 1;) {
            int n0 = r0 - 1 + row; if (n0 < 0) n0 += L0;
            __align__(16) T p[N];
            plaq_vec<T>(f, L0, L1, n0, c0 + col * N, order, p);
            sin_vec<T, N>(p);
            *reinterpret_cast<VT*>(S + row * SP + N + col * N) = *reinterpret_cast<const VT*>(p);
            col += dc; row += dr;
            if (col >= W) { col -= W; ++row; }
        }
        if (cw != L1)                                                            // the halo column c0 - 1 (whole rows: the row's own last column)
            for (int row = threadIdx.x; row < nr + 1; row += blockDim.x) {
                int n0 = r0 - 1 + row; if (n0 < 0) n0 += L0;
                S[row * SP + N - 1] = M<T>::sinv(plaq_g(f, L0, L1, n0, c0 == 0 ? L1 - 1 : c0 - 1, order));
            }
        __syncthreads();
        for (int row = row0, col = col0; row < nr;) {
            const int n1 = col * N;
            __align__(16) T s[N], su[N], f0[N], f1[N];
            *reinterpret_cast<VT*>(s) = *reinterpret_cast<const VT*>(S + (row + 1) * SP + N + n1);
            *reinterpret_cast<VT*>(su) = *reinterpret_cast<const VT*>(S + row * SP + N + n1);
            T sl = S[(row + 1) * SP + N + (n1 == 0 && cw == L1 ? L1 - 1 : n1 - 1)];
#pragma unroll
            for (int j = 0; j < N; ++j) { f0[j] = beta * (s[j] - sl); f1[j] = beta * (su[j] - s[j]); sl = s[j]; }
            *reinterpret_cast<VT*>(o + (size_t)(r0 + row) * L1 + c0 + n1) = *reinterpret_cast<const VT*>(f0);
            *reinterpret_cast<VT*>(o + (size_t)(L0 + r0 + row) * L1 + c0 + n1) = *reinterpret_cast<const VT*>(f1);
            col += dc; row += dr;
            if (col >= W) { col -= W; ++row; }
        }
    }
}

// elementwise wrap; nv 16-byte vectors, then the scalar tail
template <typename T>
__global__ void __launch_bounds__(256) k_regularize(const T* __restrict__ in, T* __restrict__ out, long long n, int vec) {
    constexpr int N = Vec<T>::N;
    using VT = typename Vec<T>::type;
    const long long stride = (long long)gridDim.x * blockDim.x, i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nv = vec ? n / N : 0;
    for (long long i = i0; i < nv; i += stride) {
        __align__(16) T v[N];
        *reinterpret_cast<VT*>(v) = reinterpret_cast<const VT*>(in)[i];
#pragma unroll
        for (int j = 0; j < N; ++j) v[j] = regularize_t(v[j]);
        reinterpret_cast<VT*>(out)[i] = *reinterpret_cast<const VT*>(v);
    }
    for (long long i = nv * N + i0; i < n; i += stride) out[i] = regularize_t(in[i]);
}

// sum of the per-(CTA, warp) gradient slices in a fixed order: out[i] = sum_s gbuf[s * n + i]
__global__ void __launch_bounds__(256) k_grad_reduce(const double* __restrict__ gbuf, int nslices, int n, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double t = 0.0;
    for (int sI = 0; sI < nslices; ++sI) t += gbuf[(size_t)sI * n + i];
    out[i] = t;
}

// fp64 FMA-pipe peak probe (roofline denominator for the resident-chain kernel; MEASURED_PEAKS.json has no
// fp64 entry).  16 independent DFMA chains per thread; 2*16*iters flop per thread.
__global__ void __launch_bounds__(256) k_dfma_probe(double* __restrict__ out, int iters) {
    double a[16];
    const double x = 1.0 + 1e-9 * threadIdx.x, y = 1e-12 * blockIdx.x;
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = 1.0 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], x, y);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 123.456) out[0] = s;      // keep the chains alive
}

// the same for the fp64 tensor path: 8 independent DMMA.8x8x4 accumulator tiles per warp; 512 flop per DMMA per warp
__global__ void __launch_bounds__(256) k_dmma_probe(double* __restrict__ out, int iters) {
    double c0[8], c1[8];
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-3 + 1e-12 * blockIdx.x;
#pragma unroll
    for (int i = 0; i < 8; ++i) { c0[i] = 1.0 + i; c1[i] = 2.0 + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c0[i] + c1[i];
    if (s == 123.456) out[0] = s;
}

#ifdef FT_PROFILE
extern "C" int fthmc_diag_profile(unsigned long long* out32_host, int reset) {
    if (out32_host) cudaMemcpyFromSymbol(out32_host, g_prof, sizeof(unsigned long long) * 32);
    if (reset) { unsigned long long z[32] = {0}; cudaMemcpyToSymbol(g_prof, z, sizeof(z)); }
    return 0;
}
#endif

extern "C" int fthmc_diag_dfma_probe(void* scratch, int iters, int blocks, void* stream, double* flop_out) {
    if (!scratch || iters <= 0 || blocks <= 0) return fail(FTHMC_E_ARG, "bad probe arguments");
    k_dfma_probe<<<blocks, 256, 0, (cudaStream_t)stream>>>((double*)scratch, iters);
    g_launches++;
    CK(cudaGetLastError());
    if (flop_out) *flop_out = 2.0 * 16.0 * (double)iters * 256.0 * (double)blocks;
    return 0;
}

extern "C" int fthmc_diag_dmma_probe(void* scratch, int iters, int blocks, void* stream, double* flop_out) {
    if (!scratch || iters <= 0 || blocks <= 0) return fail(FTHMC_E_ARG, "bad probe arguments");
    k_dmma_probe<<<blocks, 256, 0, (cudaStream_t)stream>>>((double*)scratch, iters);
    g_launches++;
    CK(cudaGetLastError());
    if (flop_out) *flop_out = 512.0 * 8.0 * (double)iters * 8.0 * (double)blocks;      // 8 tiles x 8 warps per CTA
    return 0;
}

// ------------------------------------------------------------------------------------------------
// host helpers
// ------------------------------------------------------------------------------------------------
struct fthmc_flow {
    int n_layers, act, conv, max_iter;
    double tol;
    double* wpack;   // device
    int* lmu;        // device
    int* loff;       // device
};

struct DevInfo { int sm = 0; int smem_optin = 0; bool ok = false; };
static DevInfo& devinfo() {
    static thread_local DevInfo d;
    static thread_local int dev_cached = -1;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return d;
    if (!d.ok || dev != dev_cached) {
        cudaDeviceGetAttribute(&d.sm, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        d.ok = true; dev_cached = dev;
    }
    return d;
}

static int nsm() { const int n = devinfo().sm; return n > 0 ? n : 148; }

// one thread per task (a task = one row of one stripe group; plain HMC: one site)
static int chain_threads_narrow(int L0, int L1, bool flow, int nr) {
    int tasks = (flow ? (L0 * L1) / 4 : L0 * L1) / nr;
    if (flow && FT_THREADS > 256) tasks *= 2;           // the two big convolutions split their tasks by channel pairs
    int nt = ((tasks + 31) / 32) * 32;
    return nt < 32 ? 32 : (nt > FT_THREADS ? FT_THREADS : nt);
}
// the tensor-core phases hand one 8-row tile of a stripe group to a warp: a warp per tile (the other phases leave the
// extra warps idle) shortens a chain's critical path on lattices whose task count is below 256
static int chain_threads_wide(int L0, int L1, bool flow, int nr) {
    int nt = chain_threads_narrow(L0, L1, flow, nr);
    if (flow && (L0 & 7) == 0 && (L1 & 7) == 0) {
        const int tiles = ((L0 > L1 ? L0 : L1) / 8) * ((L0 < L1 ? L0 : L1) / 4) / nr;      // max over the two orientations of G * R/8
        if (32 * tiles > nt) nt = 32 * tiles;
    }
    return nt > FT_THREADS ? FT_THREADS : nt;
}
static int chain_occupancy(int nt, size_t smem, bool flow);
// Threads per CTA for a batch of B chains.  Wide blocks cost registers, i.e. co-resident CTAs: take them when that costs
// nothing (shared memory already limits the SM to as few CTAs) or when the whole batch is resident at once anyway
// (latency matters, not throughput).
static int chain_threads(int L0, int L1, bool flow, int nr, int B);
static size_t chain_smem_bytes(int L0, int L1, bool flow, int nr) { return (engine_smem_doubles(L0, L1, flow, nr) + FT_SMEM_PREFIX) * sizeof(double); }

// static shared memory of the chain kernels (the engine object), which counts against the per-block opt-in limit
static int chain_static_smem(bool cluster) {
    static int v[2] = { -1, -1 };
    if (v[cluster] < 0) {
        cudaFuncAttributes fa;
        cudaError_t e = cluster ? cudaFuncGetAttributes(&fa, k_chain_cluster) : cudaFuncGetAttributes(&fa, k_chain);
        v[cluster] = e == cudaSuccess ? (int)fa.sharedSizeBytes : 1024;
    }
    return v[cluster];
}

// Ranks (CTAs of one thread-block cluster) a chain is spread over: 1 when the lattice fits one SM's shared memory,
// otherwise the smallest cluster size <= 16 that divides the stripe groups of both orientations and fits.  0: none.
static int chain_ranks(int L0, int L1, bool flow) {
    DevInfo& d = devinfo();
    for (int nr = 1; nr <= 16; ++nr) {
        if ((L0 / 4) % nr || (L1 / 4) % nr) continue;
        if (flow && (size_t)L0 * L1 / nr > (size_t)OFF_W3T) continue;     // Pbar plane aliases the forward weights
        // cluster mode: the link / Pbar staging areas sit in arena C behind the conv halos (chain_engine.cuh: oStage)
        if (flow && nr > 1 && 17 * (size_t)(L0 > L1 ? L0 : L1) > 4 * ((size_t)L0 * L1 / nr) + 32) continue;
        if (chain_smem_bytes(L0, L1, flow, nr) + chain_static_smem(nr > 1) > (size_t)d.smem_optin) continue;
        return nr;
    }
    return 0;
}

// number of chains (CTAs, or clusters) resident at once on the device
static int chain_occupancy(int nt, size_t smem, bool flow) {
    DevInfo& d = devinfo();
    int occ = 1;
    auto kern = flow ? k_chain : k_chain_plain;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, d.smem_optin - chain_static_smem(false));
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nt, smem) != cudaSuccess || occ < 1) occ = 1;
    return occ;
}
static int chain_threads(int L0, int L1, bool flow, int nr, int B) {
    const int n0 = chain_threads_narrow(L0, L1, flow, nr), n1 = chain_threads_wide(L0, L1, flow, nr);
    if (n1 == n0 || nr != 1) return n1;
    DevInfo& d = devinfo();
    if (!d.ok) return n0;
    const size_t smem = chain_smem_bytes(L0, L1, flow, nr);
    const int o0 = chain_occupancy(n0, smem, flow), o1 = chain_occupancy(n1, smem, flow);
    return (o1 >= o0 || (long long)B <= (long long)d.sm * o1) ? n1 : n0;
}
static int chain_resident(int L0, int L1, bool flow, int nr, int B) {
    DevInfo& d = devinfo();
    const size_t smem = chain_smem_bytes(L0, L1, flow, nr);
    const int nt = chain_threads(L0, L1, flow, nr, B);
    if (nr == 1) return d.sm * chain_occupancy(nt, smem, flow);
    cudaFuncSetAttribute(k_chain_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, d.smem_optin - chain_static_smem(true));
    cudaFuncSetAttribute(k_chain_cluster, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nr); cfg.blockDim = dim3(nt); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = nr; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int ncl = 0;
    if (cudaOccupancyMaxActiveClusters(&ncl, k_chain_cluster, &cfg) != cudaSuccess) { cudaGetLastError(); ncl = 0; }
    return ncl;
}

static int check_lattice(int B, int L0, int L1, bool flow, int* nr_out) {
    if (B <= 0 || L0 <= 0 || L1 <= 0) return fail(FTHMC_E_ARG, "B, L0, L1 must be positive");
    if (L0 % 4 || L1 % 4) return fail(FTHMC_E_LATTICE, "L0 and L1 must be multiples of 4 (4-periodic stripe masks)");
    DevInfo& d = devinfo();
    if (!d.ok) return fail(FTHMC_E_ARG, "no CUDA device");
    const int nr = chain_ranks(L0, L1, flow);
    if (nr == 0)
        return fail(FTHMC_E_LATTICE, "lattice too large for the shared-memory-resident chain path (one SM, or a cluster of up to 16 SMs) on this device");
    *nr_out = nr;
    return 0;
}

// chains resident at once on this device for (lattice, flow): what the persistent grid will be; falls back to an upper
// bound when no device can be queried
static long long resident_bound(int L0, int L1, bool flow, int nr, int B) {
    DevInfo& d = devinfo();
    if (d.ok) {
        const int r = chain_resident(L0, L1, flow, nr, B);
        if (r > 0) return r;
    }
    return nr == 1 ? (long long)nsm() * 32 : (long long)nsm() / nr;
}

extern "C" int fthmc_chain_ranks(int L0, int L1, int with_flow) {
    if (L0 <= 0 || L1 <= 0 || L0 % 4 || L1 % 4 || !devinfo().ok) return 0;
    return chain_ranks(L0, L1, with_flow != 0);
}

extern "C" size_t fthmc_workspace_bytes(fthmc_flow_t flow, int B, int L0, int L1) {
    if (B <= 0 || L0 <= 0 || L1 <= 0 || L0 % 4 || L1 % 4) return 0;
    int nr = chain_ranks(L0, L1, flow != nullptr);
    if (nr == 0) nr = 1;
    long long g = resident_bound(L0, L1, flow != nullptr, nr, B);
    if (B < g) g = B;
    if (g < 1) g = 1;
    return (size_t)g * engine_ws_doubles(L0, L1, flow ? flow->n_layers : 0, nr) * sizeof(double) + 256;
}

static int launch_chain(ChainArgs& a, fthmc_flow_t flow, int L0, int L1, void* ws, size_t ws_bytes, void* stream) {
    const bool has_flow = flow != nullptr;
    int nr = 1;
    int rc = check_lattice(a.B, L0, L1, has_flow, &nr);
    if (rc) return rc;
    a.pr.L0 = L0; a.pr.L1 = L1;
    if (has_flow) {
        a.pr.nlayers = flow->n_layers; a.pr.act = flow->act; a.pr.conv = flow->conv;
        a.pr.inv_tol = flow->tol; a.pr.inv_max_iter = flow->max_iter;
        a.pr.wpack = flow->wpack; a.pr.lmu = flow->lmu; a.pr.loff = flow->loff;
    } else {
        a.pr.nlayers = 0; a.pr.act = 0; a.pr.conv = 0; a.pr.inv_tol = 0; a.pr.inv_max_iter = 0;
        a.pr.wpack = nullptr; a.pr.lmu = nullptr; a.pr.loff = nullptr;
    }
    const bool train = a.mode == MODE_FT_GRAD;
    a.pr.train = train ? 1 : 0;
    if (train && (nr != 1 || (L0 & 7) || (L1 & 7)))
        return fail(FTHMC_E_LATTICE, "the weight-gradient path needs L0, L1 multiples of 8 and a lattice that fits one SM (L0*L1 <= 1024)");
    int res = chain_resident(L0, L1, has_flow, nr, a.B);
    if (res < 1) return fail(FTHMC_E_LATTICE, "the device cannot co-schedule a thread-block cluster of the size this lattice needs");
    const int chains = a.B < res ? a.B : res;
    a.ws_stride = engine_ws_doubles(L0, L1, a.pr.nlayers, nr, train);
    const int nt = chain_threads(L0, L1, has_flow, nr, a.B);
    size_t need = (size_t)chains * a.ws_stride * sizeof(double);
    if (train) {            // gradient accumulators behind the chain workspaces: [CTA][warp][layer][GRAD_DOUBLES]
        a.gbuf_stride = (size_t)(nt / 32) * a.pr.nlayers * GRAD_DOUBLES;
        a.gbuf = (double*)ws + (size_t)chains * a.ws_stride;
        need += (size_t)chains * a.gbuf_stride * sizeof(double);
    }
    if (!ws || ws_bytes < need) return fail(FTHMC_E_WORKSPACE, "workspace null or smaller than fthmc_workspace_bytes() / fthmc_grad_workspace_bytes()");
    if (((uintptr_t)ws) & 15) return fail(FTHMC_E_WORKSPACE, "workspace must be 16-byte aligned");
    a.ws = (double*)ws;
    if (train) CK(cudaMemsetAsync(a.gbuf, 0, (size_t)chains * a.gbuf_stride * sizeof(double), (cudaStream_t)stream));
    const size_t smem = chain_smem_bytes(L0, L1, has_flow, nr);
    // Flow programs run one device wave (the co-resident chains) per launch, back to back on the stream: the per-wave time
    // of a long persistent launch creeps up as the SMs drift apart (6.70 ms for one wave, 6.90 ms per wave over 28 at
    // L = 32, profiles/r1_wave_scaling.txt; 4096 chains: 192.9 ms in one launch, 186.2 ms in 28, profiles/r1_chunk_scaling.txt),
    // and a launch boundary realigns them for a few microseconds.  Chains are addressed by their global index, so the
    // split is invisible in the results.  The weight-gradient mode keeps one launch (its per-CTA accumulators are
    // reduced after the launch), and so does plain HMC (a wave there lasts tens of microseconds).
    const int per_launch = (has_flow && !train && FT_WAVE_LAUNCHES) ? chains : a.B;
    for (int b0 = 0; b0 < a.B; b0 += per_launch) {
        a.b_begin = b0; a.b_end = b0 + per_launch < a.B ? b0 + per_launch : a.B;
        const int grid = a.b_end - a.b_begin < chains ? a.b_end - a.b_begin : chains;
        if (nr == 1) {
            if (has_flow) k_chain<<<grid, nt, smem, (cudaStream_t)stream>>>(a);
            else k_chain_plain<<<grid, nt, smem, (cudaStream_t)stream>>>(a);
        } else {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(grid * nr); cfg.blockDim = dim3(nt); cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = nr; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            CK(cudaLaunchKernelEx(&cfg, k_chain_cluster, a));
        }
        g_launches++;
    }
    CK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// C ABI: stencils
// ------------------------------------------------------------------------------------------------
// CTAs per chain of the reduction stencils: 1 when the batch alone fills the device, otherwise a cluster of up to 16
template <typename T, int WHAT, int ORDER>
static int reduce_launch_t(const void* links, int B, int L0, int L1, double beta, int rounded, void* out, cudaStream_t st) {
    int nc = 1;
    while (nc < 16 && (long long)B * nc < 4 * nsm() && 2 * nc <= L0 && (long long)(L0 / (2 * nc)) * L1 >= 1024) nc *= 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nc, B); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = nc; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    // 16-byte vector path: whole vectors per row, 16-byte aligned rows
    const bool vec = L1 % Vec<T>::N == 0 && ((uintptr_t)links & 15) == 0;
    const size_t chain_bytes = (size_t)2 * L0 * L1 * sizeof(T);
    if (FT_STENCIL_TMA && nc == 1 && vec && chain_bytes <= 32 * 1024 && chain_bytes % 16 == 0 && L0 * (L1 / Vec<T>::N) >= 64 && B >= 16 * nsm()) {
        constexpr int NSTAGE = FT_TMA_STAGES;
        const int stage_elems = (int)(((chain_bytes + 127) / 128) * 128 / sizeof(T));
        const size_t smem = (size_t)NSTAGE * stage_elems * sizeof(T) + NSTAGE * 8 + 16 * 8;
        auto kern = k_action_topo_tma<T, NSTAGE, WHAT, ORDER>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // persistent CTAs: exactly as many as are resident at once (registers or shared memory, whichever binds)
        int per_sm = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));
        if (per_sm < 1) per_sm = 1;
        if ((size_t)per_sm * (smem + 1024) > 200 * 1024) per_sm = (int)((200 * 1024) / (smem + 1024));
        int grid = nsm() * per_sm; if (grid > B) grid = B;
        kern<<<grid, 256, smem, st>>>((const T*)links, B, L0, L1, stage_elems, beta, rounded, (T*)out);
        g_launches += 1;
        CK(cudaGetLastError());
        return 0;
    }
    if (nc == 1 && vec && (long long)L0 * L1 <= 4096 && L0 * (L1 / Vec<T>::N) >= 32 && B >= 16 * nsm()) {
        k_action_topo_warp<T, WHAT, ORDER><<<(B + 7) / 8, 256, 0, st>>>((const T*)links, B, L0, L1, beta, rounded, (T*)out);
        g_launches += 1;
        CK(cudaGetLastError());
        return 0;
    }
    // (fp32 scans measured 5 % slower pipelined: four sites per vector already give their loop twice the bytes in flight)
    auto kern = vec ? (nc > 1 && sizeof(T) == 8 ? k_action_topo<T, true, WHAT, ORDER, true> : k_action_topo<T, true, WHAT, ORDER>) : k_action_topo<T, false, WHAT, ORDER>;
    if (nc > 8) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    // all clusters in one wave: a 16-CTA cluster needs 16 free slots inside ONE GPC, and fewer of those exist than the
    // CTA slots of the device suggest (45 at once on a B200, 104 clusters of 8) -- three stragglers of 48 chains cost more
    // than the lower occupancy of the next smaller cluster size (L = 1024, 48 chains: +4 %)
    static int max_active[2][5] = { { 0, 0, 0, 0, 0 }, { 0, 0, 0, 0, 0 } };   // [vec][log2 nc], this instantiation's kernels
    while (nc > 1) {
        int lg = 0; while ((1 << lg) < nc) ++lg;
        int& mc = max_active[vec ? 1 : 0][lg];
        if (mc == 0) {
            cfg.gridDim = dim3(nc, B); at[0].val.clusterDim.x = nc;
            if (cudaOccupancyMaxActiveClusters(&mc, kern, &cfg) != cudaSuccess || mc < 1) { mc = 1 << 30; (void)cudaGetLastError(); }
        }
        if (B <= mc) break;
        nc /= 2;
    }
    const int rows = (L0 + nc - 1) / nc;
    cfg.gridDim = dim3(nc, B); at[0].val.clusterDim.x = nc;
    CK(cudaLaunchKernelEx(&cfg, kern, (const T*)links, L0, L1, rows, beta, rounded, (T*)out));
    g_launches += 1;
    CK(cudaGetLastError());
    return 0;
}
// (what, order) pairs in use: action with either plaquette term order, floored charge (order 1), batched charge (order 0)
template <typename T>
static int reduce_launch(const void* links, int B, int L0, int L1, int what, int order, double beta, int rounded, void* out, cudaStream_t st) {
    if (what == 0) return order == 0 ? reduce_launch_t<T, 0, 0>(links, B, L0, L1, beta, rounded, out, st)
                                     : reduce_launch_t<T, 0, 1>(links, B, L0, L1, beta, rounded, out, st);
    if (what == 1) return reduce_launch_t<T, 1, 1>(links, B, L0, L1, beta, rounded, out, st);
    return reduce_launch_t<T, 2, 0>(links, B, L0, L1, beta, rounded, out, st);
}

static int check_stencil(const void* in, const void* out, int B, int L0, int L1, int dtype) {
    if (!in || !out) return fail(FTHMC_E_ARG, "null pointer");
    if (B <= 0 || L0 <= 0 || L1 <= 0) return fail(FTHMC_E_ARG, "B, L0, L1 must be positive");
    if (dtype != FTHMC_F64 && dtype != FTHMC_F32) return fail(FTHMC_E_DTYPE, "dtype must be FTHMC_F64 or FTHMC_F32");
    return 0;
}

// the chain-per-CTA stencils carry the batch in grid.y (<= 65535): larger batches go in consecutive launches
constexpr int STENCIL_MAX_B = 65535;
template <class F> static int for_batch_chunks(int B, F launch) {
    for (int b0 = 0; b0 < B; b0 += STENCIL_MAX_B) {
        const int rc = launch(b0, B - b0 < STENCIL_MAX_B ? B - b0 : STENCIL_MAX_B);
        if (rc) return rc;
    }
    return 0;
}

extern "C" int fthmc_action(const void* links, int B, int L0, int L1, double beta, int order, void* out, int dtype, void* stream) {
    int rc = check_stencil(links, out, B, L0, L1, dtype); if (rc) return rc;
    if (order != 0 && order != 1) return fail(FTHMC_E_ARG, "order must be 0 or 1");
    const size_t es = dtype == FTHMC_F64 ? 8 : 4, cs = (size_t)2 * L0 * L1 * es;
    return for_batch_chunks(B, [&](int b0, int nb) {
        const void* in = (const char*)links + b0 * cs; void* o = (char*)out + b0 * es;
        return dtype == FTHMC_F64 ? reduce_launch<double>(in, nb, L0, L1, 0, order, beta, 0, o, (cudaStream_t)stream)
                                  : reduce_launch<float>(in, nb, L0, L1, 0, order, beta, 0, o, (cudaStream_t)stream);
    });
}

extern "C" int fthmc_topo_charge(const void* links, int B, int L0, int L1, int rounded, void* out, int dtype, void* stream) {
    int rc = check_stencil(links, out, B, L0, L1, dtype); if (rc) return rc;
    // rounded: hmc_2dU1.topocharge (plaqphase order, regularize); else field_transformation.topo_charge (u1_plaq order, torch_wrap)
    const int what = rounded ? 1 : 2, order = rounded ? 1 : 0;
    const size_t es = dtype == FTHMC_F64 ? 8 : 4, cs = (size_t)2 * L0 * L1 * es;
    return for_batch_chunks(B, [&](int b0, int nb) {
        const void* in = (const char*)links + b0 * cs; void* o = (char*)out + b0 * es;
        return dtype == FTHMC_F64 ? reduce_launch<double>(in, nb, L0, L1, what, order, 0.0, rounded, o, (cudaStream_t)stream)
                                  : reduce_launch<float>(in, nb, L0, L1, what, order, 0.0, rounded, o, (cudaStream_t)stream);
    });
}

extern "C" int fthmc_force(const void* links, int B, int L0, int L1, double beta, int order, void* force_out, int dtype, void* stream) {
    int rc = check_stencil(links, force_out, B, L0, L1, dtype); if (rc) return rc;
    if (order != 0 && order != 1) return fail(FTHMC_E_ARG, "order must be 0 or 1");
    const size_t es = dtype == FTHMC_F64 ? 8 : 4;
    const int vw = dtype == FTHMC_F64 ? Vec<double>::N : Vec<float>::N;
    const bool vec = L1 % vw == 0 && ((uintptr_t)links & 15) == 0 && ((uintptr_t)force_out & 15) == 0;
    // tile: about 24 KB of sin(P) per CTA keeps eight CTAs (all 2048 threads) on an SM, so that the load / sine half of one
    // tile overlaps the store half of others.  Rows of 4 KB and more are cut into 1 KB column chunks (vector path), so that
    // a tile still has ~22 rows per halo row; without the vector path very wide rows take what 64 KB hold.
    int cw = L1;
    if (vec && (size_t)L1 * es >= 4096 && L1 % (int)(1024 / es) == 0) cw = (int)(1024 / es);
    const size_t pitch = (size_t)(cw + (cw != L1 ? vw : 0)) * es;
    const int rows_max = (int)((64 * 1024) / pitch) - 1;
    if (rows_max < 1) return fail(FTHMC_E_LATTICE, "L1 too large for the force tile");
    int rows = (int)((24 * 1024) / pitch) - 1;
    if (rows < 15) rows = rows_max < 15 ? rows_max : 15;
    if (rows > L0) rows = L0;
    rows = (L0 + (L0 + rows - 1) / rows - 1) / ((L0 + rows - 1) / rows);          // equal tiles
    // small batches of large lattices: shorter tiles until the grid fills the device
    while (rows > 8 && (long long)B * ((L0 + rows - 1) / rows) * (L1 / cw) < 4 * nsm()) rows = (rows + 1) / 2;
    const int nchunk = ((L0 + rows - 1) / rows) * (L1 / cw);
    const size_t smem = (size_t)(rows + 1) * pitch;
    const size_t cs = (size_t)2 * L0 * L1 * es;
    return for_batch_chunks(B, [&](int b0, int nb) -> int {
        const char* in = (const char*)links + b0 * cs; char* fo = (char*)force_out + b0 * cs;
        if (dtype == FTHMC_F64) {
            if (cw != L1) k_force_tiled<double><<<dim3(nchunk, nb), 256, smem, (cudaStream_t)stream>>>((const double*)in, L0, L1, rows, cw, beta, order, (double*)fo);
            else {
                auto kern = vec ? k_force<double, true> : k_force<double, false>;
                if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
                kern<<<dim3(nchunk, nb), 256, smem, (cudaStream_t)stream>>>((const double*)in, L0, L1, rows, beta, order, (double*)fo);
            }
        } else {
            if (cw != L1) k_force_tiled<float><<<dim3(nchunk, nb), 256, smem, (cudaStream_t)stream>>>((const float*)in, L0, L1, rows, cw, (float)beta, order, (float*)fo);
            else {
                auto kern = vec ? k_force<float, true> : k_force<float, false>;
                if (smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
                kern<<<dim3(nchunk, nb), 256, smem, (cudaStream_t)stream>>>((const float*)in, L0, L1, rows, (float)beta, order, (float*)fo);
            }
        }
        g_launches++;
        CK(cudaGetLastError());
        return 0;
    });
}

extern "C" int fthmc_regularize(const void* in, void* out, long long n, int dtype, void* stream) {
    if (!in || !out || n <= 0) return fail(FTHMC_E_ARG, "null pointer or n <= 0");
    const int vec = (((uintptr_t)in | (uintptr_t)out) & 15) == 0;
    long long blocks = (n / 2 + 255) / 256; if (blocks < 1) blocks = 1; if (blocks > nsm() * 32) blocks = nsm() * 32;
    if (dtype == FTHMC_F64) k_regularize<double><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const double*)in, (double*)out, n, vec);
    else if (dtype == FTHMC_F32) k_regularize<float><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const float*)in, (float*)out, n, vec);
    else return fail(FTHMC_E_DTYPE, "dtype must be FTHMC_F64 or FTHMC_F32");
    g_launches++;
    CK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// C ABI: flow handle
// ------------------------------------------------------------------------------------------------
extern "C" int fthmc_flow_pack(const double* raw_host, int n_layers, const int* mu_host, const int* off_host,
                               int hidden0, int hidden1, int n_mix, int ksize, int activation, int convention,
                               double inv_tol, int inv_max_iter, fthmc_flow_t* out) {
    if (!raw_host || !mu_host || !off_host || !out || n_layers <= 0) return fail(FTHMC_E_ARG, "null pointer or n_layers <= 0");
    if (n_layers > 128) return fail(FTHMC_E_ARG, "at most 128 coupling layers");
    if (hidden0 != NH || hidden1 != NH || n_mix != NK || ksize != 3)
        return fail(FTHMC_E_NETSHAPE, "only the reference CNN shape is built: hidden_sizes=[8,8], n_mixture_comps=2, kernel_size=3");
    if (activation < 0 || activation > 2) return fail(FTHMC_E_ARG, "activation must be silu(0), leaky_relu(1) or relu(2)");
    if (convention != 0 && convention != 1) return fail(FTHMC_E_ARG, "convention must be 0 ([0,2pi)) or 1 ([-pi,pi))");
    if (!(inv_tol > 0) || inv_max_iter <= 0) return fail(FTHMC_E_ARG, "inv_tol and inv_max_iter must be positive");
    for (int l = 0; l < n_layers; ++l)
        if ((mu_host[l] != 0 && mu_host[l] != 1) || off_host[l] < 0 || off_host[l] > 3)
            return fail(FTHMC_E_ARG, "mask mu must be 0/1 and mask off in 0..3");
    std::vector<double> pack((size_t)n_layers * PACK_DOUBLES);
    for (int l = 0; l < n_layers; ++l) pack_layer(raw_host + (size_t)l * RAW_DOUBLES, mu_host[l], pack.data() + (size_t)l * PACK_DOUBLES);
    fthmc_flow* f = new fthmc_flow();
    f->n_layers = n_layers; f->act = activation; f->conv = convention; f->tol = inv_tol; f->max_iter = inv_max_iter;
    f->wpack = nullptr; f->lmu = nullptr; f->loff = nullptr;
    cudaError_t e;
    if ((e = cudaMalloc(&f->wpack, pack.size() * sizeof(double))) != cudaSuccess ||
        (e = cudaMalloc(&f->lmu, n_layers * sizeof(int))) != cudaSuccess ||
        (e = cudaMalloc(&f->loff, n_layers * sizeof(int))) != cudaSuccess ||
        (e = cudaMemcpy(f->wpack, pack.data(), pack.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(f->lmu, mu_host, n_layers * sizeof(int), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(f->loff, off_host, n_layers * sizeof(int), cudaMemcpyHostToDevice)) != cudaSuccess) {
        cudaFree(f->wpack); cudaFree(f->lmu); cudaFree(f->loff); delete f;
        return cuda_fail(e, "fthmc_flow_pack");
    }
    *out = f;
    return 0;
}

// new weights into an existing handle (same layer count and masks): the training loop's per-step re-pack without
// re-allocating device memory.  Synchronous (it uploads); no launch using the handle may be in flight.
extern "C" int fthmc_flow_update(fthmc_flow_t f, const double* raw_host) {
    if (!f || !raw_host) return fail(FTHMC_E_ARG, "null pointer");
    std::vector<int> mu(f->n_layers);
    CK(cudaMemcpy(mu.data(), f->lmu, f->n_layers * sizeof(int), cudaMemcpyDeviceToHost));
    std::vector<double> pack((size_t)f->n_layers * PACK_DOUBLES);
    for (int l = 0; l < f->n_layers; ++l) pack_layer(raw_host + (size_t)l * RAW_DOUBLES, mu[l], pack.data() + (size_t)l * PACK_DOUBLES);
    CK(cudaMemcpy(f->wpack, pack.data(), pack.size() * sizeof(double), cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int fthmc_flow_free(fthmc_flow_t f) {
    if (!f) return 0;
    cudaFree(f->wpack); cudaFree(f->lmu); cudaFree(f->loff);
    delete f;
    return 0;
}

extern "C" int fthmc_flow_n_layers(fthmc_flow_t f) { return f ? f->n_layers : 0; }

// ------------------------------------------------------------------------------------------------
// C ABI: resident-chain entry points
// ------------------------------------------------------------------------------------------------
extern "C" int fthmc_flow_fwd(fthmc_flow_t flow, const double* x_in, double* x_out, double* logJ, double* layer_logJ,
                              int B, int L0, int L1, void* ws, size_t ws_bytes, void* stream) {
    if (!flow || !x_in || !x_out) return fail(FTHMC_E_ARG, "null pointer");
    ChainArgs a{}; a.mode = MODE_FLOW_FWD; a.B = B; a.field_in = x_in; a.field_out = x_out; a.s_out = logJ; a.layer_logJ = layer_logJ;
    return launch_chain(a, flow, L0, L1, ws, ws_bytes, stream);
}

extern "C" int fthmc_flow_inv(fthmc_flow_t flow, const double* x_in, double* x_out, double* logJ, double* layer_logJ, int* iters,
                              int B, int L0, int L1, void* ws, size_t ws_bytes, void* stream) {
    if (!flow || !x_in || !x_out) return fail(FTHMC_E_ARG, "null pointer");
    ChainArgs a{}; a.mode = MODE_FLOW_INV; a.B = B; a.field_in = x_in; a.field_out = x_out; a.s_out = logJ; a.layer_logJ = layer_logJ;
    a.iters = iters;
    return launch_chain(a, flow, L0, L1, ws, ws_bytes, stream);
}

extern "C" int fthmc_ft_action(fthmc_flow_t flow, const double* x, double beta, double* out, double* flowed,
                               int B, int L0, int L1, void* ws, size_t ws_bytes, void* stream) {
    if (!flow || !x || !out) return fail(FTHMC_E_ARG, "null pointer");
    ChainArgs a{}; a.mode = MODE_FT_ACTION; a.B = B; a.field_in = x; a.field_out = flowed; a.s_out = out; a.beta = beta;
    return launch_chain(a, flow, L0, L1, ws, ws_bytes, stream);
}

extern "C" int fthmc_ft_force(fthmc_flow_t flow, const double* x, double beta, double* force_out,
                              int B, int L0, int L1, void* ws, size_t ws_bytes, void* stream) {
    if (!flow || !x || !force_out) return fail(FTHMC_E_ARG, "null pointer");
    ChainArgs a{}; a.mode = MODE_FT_FORCE; a.B = B; a.field_in = x; a.field_out = force_out; a.beta = beta;
    return launch_chain(a, flow, L0, L1, ws, ws_bytes, stream);
}

static int leapfrog_common(fthmc_flow_t flow, int mode, const double* x_in, const double* p_in, double* x_out, double* p_out,
                           int B, int L0, int L1, double beta, double dt, int nstep, void* ws, size_t ws_bytes, void* stream) {
    if (!x_in || !p_in || !x_out || !p_out) return fail(FTHMC_E_ARG, "null pointer");
    if (nstep < 1) return fail(FTHMC_E_ARG, "nstep must be >= 1");
    ChainArgs a{}; a.mode = mode; a.B = B; a.field_in = x_in; a.p_in = p_in; a.field_out = x_out; a.p_out = p_out;
    a.beta = beta; a.dt = dt; a.nstep = nstep;
    return launch_chain(a, flow, L0, L1, ws, ws_bytes, stream);
}

extern "C" int fthmc_leapfrog(const double* x_in, const double* p_in, double* x_out, double* p_out, int B, int L0, int L1,
                              double beta, double dt, int nstep, void* ws, size_t ws_bytes, void* stream) {
    return leapfrog_common(nullptr, MODE_LEAPFROG, x_in, p_in, x_out, p_out, B, L0, L1, beta, dt, nstep, ws, ws_bytes, stream);
}

extern "C" int fthmc_ft_leapfrog(fthmc_flow_t flow, const double* x_in, const double* p_in, double* x_out, double* p_out,
                                 int B, int L0, int L1, double beta, double dt, int nstep, void* ws, size_t ws_bytes, void* stream) {
    if (!flow) return fail(FTHMC_E_ARG, "null flow");
    return leapfrog_common(flow, MODE_FT_LEAPFROG, x_in, p_in, x_out, p_out, B, L0, L1, beta, dt, nstep, ws, ws_bytes, stream);
}

static int traj_common(fthmc_flow_t flow, int mode, const double* field_in, double* field_out, const double* p_in, const double* u_in,
                       unsigned long long seed, unsigned long long traj, unsigned long long chain0,
                       int B, int L0, int L1, double beta, double dt, int nstep,
                       double* dH, double* exp_mdH, int* acc, double* plaq, double* topo, double* h0, double* h1,
                       void* ws, size_t ws_bytes, void* stream, int ntraj = 1) {
    if (!field_in || !field_out) return fail(FTHMC_E_ARG, "null pointer");
    if (nstep < 1) return fail(FTHMC_E_ARG, "nstep must be >= 1");
    if (ntraj < 1) return fail(FTHMC_E_ARG, "ntraj must be >= 1");
    ChainArgs a{}; a.mode = mode; a.B = B; a.field_in = field_in; a.field_out = field_out; a.p_in = p_in; a.u_in = u_in;
    a.ntraj = ntraj;
    a.seed = seed; a.traj = traj; a.chain0 = chain0; a.beta = beta; a.dt = dt; a.nstep = nstep;
    a.s_out = dH; a.expmdH = exp_mdH; a.acc = acc; a.plaq = plaq; a.topo = topo; a.h0 = h0; a.h1 = h1;
    return launch_chain(a, flow, L0, L1, ws, ws_bytes, stream);
}

extern "C" int fthmc_hmc_traj(const double* x_in, double* x_out, const double* p_in, const double* u_in,
                              unsigned long long seed, unsigned long long traj, unsigned long long chain0,
                              int B, int L0, int L1, double beta, double dt, int nstep,
                              double* dH, double* exp_mdH, int* acc, double* plaq, double* topo,
                              void* ws, size_t ws_bytes, void* stream) {
    return traj_common(nullptr, MODE_HMC, x_in, x_out, p_in, u_in, seed, traj, chain0, B, L0, L1, beta, dt, nstep,
                       dH, exp_mdH, acc, plaq, topo, nullptr, nullptr, ws, ws_bytes, stream);
}

extern "C" int fthmc_ft_hmc_traj(fthmc_flow_t flow, const double* field_in, double* field_out, const double* p_in, const double* u_in,
                                 unsigned long long seed, unsigned long long traj, unsigned long long chain0,
                                 int B, int L0, int L1, double beta, double dt, int nstep,
                                 double* dH, double* exp_mdH, int* acc, double* plaq, double* topo, double* h0, double* h1,
                                 void* ws, size_t ws_bytes, void* stream) {
    if (!flow) return fail(FTHMC_E_ARG, "null flow");
    return traj_common(flow, MODE_FT_HMC, field_in, field_out, p_in, u_in, seed, traj, chain0, B, L0, L1, beta, dt, nstep,
                       dH, exp_mdH, acc, plaq, topo, h0, h1, ws, ws_bytes, stream);
}

// ------------------------------------------------------------------------------------------------
// C ABI: run loops -- ntraj consecutive trajectories per chain in ONE launch, the field resident in shared memory
// ------------------------------------------------------------------------------------------------
extern "C" int fthmc_hmc_run(const double* x_in, double* x_out, const double* p_in, const double* u_in,
                             unsigned long long seed, unsigned long long traj0, unsigned long long chain0,
                             int B, int L0, int L1, double beta, double dt, int nstep, int ntraj,
                             double* dH, double* exp_mdH, int* acc, double* plaq, double* topo,
                             void* ws, size_t ws_bytes, void* stream) {
    return traj_common(nullptr, MODE_HMC, x_in, x_out, p_in, u_in, seed, traj0, chain0, B, L0, L1, beta, dt, nstep,
                       dH, exp_mdH, acc, plaq, topo, nullptr, nullptr, ws, ws_bytes, stream, ntraj);
}

extern "C" int fthmc_ft_hmc_run(fthmc_flow_t flow, const double* field_in, double* field_out, const double* p_in, const double* u_in,
                                unsigned long long seed, unsigned long long traj0, unsigned long long chain0,
                                int B, int L0, int L1, double beta, double dt, int nstep, int ntraj,
                                double* dH, double* exp_mdH, int* acc, double* plaq, double* topo,
                                void* ws, size_t ws_bytes, void* stream) {
    if (!flow) return fail(FTHMC_E_ARG, "null flow");
    return traj_common(flow, MODE_FT_HMC, field_in, field_out, p_in, u_in, seed, traj0, chain0, B, L0, L1, beta, dt, nstep,
                       dH, exp_mdH, acc, plaq, topo, nullptr, nullptr, ws, ws_bytes, stream, ntraj);
}

// ------------------------------------------------------------------------------------------------
// C ABI: flow-training gradient
// ------------------------------------------------------------------------------------------------
extern "C" size_t fthmc_grad_workspace_bytes(fthmc_flow_t flow, int B, int L0, int L1) {
    if (!flow || B <= 0 || L0 <= 0 || L1 <= 0 || L0 % 4 || L1 % 4) return 0;
    long long g = resident_bound(L0, L1, true, 1, B);
    if (B < g) g = B;
    if (g < 1) g = 1;
    return (size_t)g * (engine_ws_doubles(L0, L1, flow->n_layers, 1, true) + (size_t)(FT_THREADS / 32) * flow->n_layers * GRAD_DOUBLES)
               * sizeof(double) + 256;
}

extern "C" int fthmc_ft_action_grad(fthmc_flow_t flow, const double* x, double beta, double* action_out, double* grad_canon, double* force_out,
                                    int B, int L0, int L1, void* ws, size_t ws_bytes, void* stream) {
    if (!flow || !x || !grad_canon) return fail(FTHMC_E_ARG, "null pointer");
    ChainArgs a{}; a.mode = MODE_FT_GRAD; a.B = B; a.field_in = x; a.field_out = force_out; a.s_out = action_out; a.beta = beta;
    int rc = launch_chain(a, flow, L0, L1, ws, ws_bytes, stream);
    if (rc) return rc;
    const int n = flow->n_layers * GRAD_DOUBLES;
    const int chains = (int)(((char*)a.gbuf - (char*)a.ws) / (a.ws_stride * sizeof(double)));
    const int nslices = chains * (int)(a.gbuf_stride / ((size_t)flow->n_layers * GRAD_DOUBLES));
    k_grad_reduce<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a.gbuf, nslices, n, grad_canon);
    g_launches++;
    CK(cudaGetLastError());
    return 0;
}

// vector-Jacobian product of the flow (x, weights) -> (y, logJ):  d/dx and d/dweights of  sum_b [ <gy_b, y_b> + glj_b logJ_b ]
extern "C" int fthmc_flow_vjp(fthmc_flow_t flow, const double* x, const double* gy, const double* glj, double* grad_canon, double* grad_x,
                              int B, int L0, int L1, void* ws, size_t ws_bytes, void* stream) {
    if (!flow || !x || !gy || !glj || !grad_canon) return fail(FTHMC_E_ARG, "null pointer");
    ChainArgs a{}; a.mode = MODE_FT_GRAD; a.B = B; a.field_in = x; a.field_out = grad_x; a.s_out = nullptr; a.beta = 0.0;
    a.vjp_seed = gy; a.vjp_wlj = glj;
    int rc = launch_chain(a, flow, L0, L1, ws, ws_bytes, stream);
    if (rc) return rc;
    const int n = flow->n_layers * GRAD_DOUBLES;
    const int chains = (int)(((char*)a.gbuf - (char*)a.ws) / (a.ws_stride * sizeof(double)));
    const int nslices = chains * (int)(a.gbuf_stride / ((size_t)flow->n_layers * GRAD_DOUBLES));
    k_grad_reduce<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a.gbuf, nslices, n, grad_canon);
    g_launches++;
    CK(cudaGetLastError());
    return 0;
}

extern "C" int fthmc_grad_doubles(void) { return GRAD_DOUBLES; }

extern "C" int fthmc_grad_unpack(const double* grad_canon_host, int n_layers, const int* mu_host, double* raw_host) {
    if (!grad_canon_host || !mu_host || !raw_host || n_layers <= 0) return fail(FTHMC_E_ARG, "null pointer or n_layers <= 0");
    for (int l = 0; l < n_layers; ++l)
        unpack_grad_layer(grad_canon_host + (size_t)l * GRAD_DOUBLES, mu_host[l], raw_host + (size_t)l * RAW_DOUBLES);
    return 0;
}
