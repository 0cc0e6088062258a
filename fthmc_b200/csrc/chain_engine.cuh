// chain_engine.cuh -- one Markov chain's FT-HMC state machine, resident in shared memory.
//
// One CTA (or, for lattices beyond one SM, one thread-block cluster) owns one chain.  The chain's link
// field, the gradient (force) field and every CNN activation plane live in shared memory for the whole
// call; HBM/L2 is touched only to load/store the field at the boundaries and for the per-chain workspace
// (momenta, and the layer blocks the forward sweep of ft_force leaves for its adjoint sweep).
//
// The code is written as block-cooperative *phases*: each phase is a strided loop over "tasks"
// (task = one row of one 4-column stripe group; the tensor-core phases: one 8-row tile per warp)
// followed by a barrier.  Every phase is correct for ANY thread count, including 1, and the warp-level
// phases carry their lanes in arrays (FT_LANES): that is what lets tests/emul compile this same header
// with g++ and run it on the CPU (test infrastructure only; the product is the CUDA build).
//
// Contents: scalar helpers (exp_fast, activations, Philox) | Engine: geometry and field access, Wilson
// pieces, coupling-layer phases (planes, conv1, conv2 [DFMA and DMMA forms], conv3 forward / reverse), adjoint
// phases (outgrad, conv3^T, conv2^T, conv1^T, scatter), TMA prefetch and the per-layer sweeps, weight-gradient
// phases (training) | trajectory programs (leapfrog, ft_hmc, hmc).
//
// Reference semantics (paths relative to nftqcd/fthmc):
//   coupling layer forward/reverse  ipynb/field_transformation.py:152-174, 288-338
//   masks                           ipynb/field_transformation.py:175-248
//   tan mixture + logJ + bisection  ipynb/field_transformation.py:249-285
//   ft_flow/ft_flow_inv/ft_action/ft_force   ipynb/ft_hmc.py:220-249
//   ft_leapfrog/ft_hmc              ipynb/ft_hmc.py:394-435
//   action/force/regularize/topocharge/leapfrog/hmc   hmc_2dU1.py:100-155
// The force of the flowed action, which the reference takes from torch.autograd, is the
// hand-derived adjoint of SURVEY.md section 8a row 11 (validated against autograd by
// oracle.ft_force_adjoint in tests/test_oracle_golden.py).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define FT_HD __host__ __device__ __forceinline__
#define FT_PHASE __host__ __device__ __noinline__
#else
#define FT_HD inline
#define FT_PHASE
#endif

// Optional per-phase cycle accounting (development builds only: -DFT_PROFILE).  E must then provide
// clock() and prof_add(id, cycles).
#ifdef FT_PROFILE
#define FT_T(id, ...) do { long long t0_ = ex.clock(); __VA_ARGS__; ex.prof_add(id, ex.clock() - t0_); } while (0)
#else
#define FT_T(id, ...) do { __VA_ARGS__; } while (0)
#endif

#ifndef FT_BISECT_REPLAY
#define FT_BISECT_REPLAY 1
#endif
#ifndef FT_WINOGRAD
#define FT_WINOGRAD 1
#endif
#ifndef FT_CONV1T_PAIR
#define FT_CONV1T_PAIR 1
#endif
#ifndef FT_CONV3T_PAIR
#define FT_CONV3T_PAIR 1
#endif
// act'(z2) of the adjoint sweep single-buffered in plane B, fetched one layer ahead (measured: no slower than the round-1
// double buffer B / C, profiles/r2_microopt_ab.txt) -- which leaves plane C free for the trajectory state: momenta, x0, y0.
#ifndef FT_D2_SINGLE
#define FT_D2_SINGLE 1
#endif

// loop over the lanes a "thread" carries: exactly one iteration (its own lane) on the device, 32 in the serial emulation
#define FT_LANES(ln, ls) for (int ls = 0, ln = ex.lane0(); ls < E::kLanes; ++ls, ++ln)

namespace fthmc {

enum ProfId { PF_PLANES = 0, PF_CONV1, PF_CONV2, PF_CONV3F, PF_CONV3R, PF_OUTGRAD, PF_CONV3T, PF_CONV2T, PF_CONV1T,
              PF_SCATTER, PF_ISSUE, PF_WFORCE, PF_LEAP, PF_MISC, PF_C2_MAC, PF_C2_ACT, PF_C1_MAC, PF_C1_ACT, PF_C2T_MAC, PF_C3_CONV, PF_N };

// ---- network shape (every reference config: hidden_sizes=[8,8], n_mixture_comps=2, kernel 3) ----
constexpr int NH = 8;          // hidden channels of both hidden layers
constexpr int NK = 2;          // mixture components (s_1..s_K), +1 channel for t
constexpr int NOUT = NK + 1;

// ---- packed per-layer weight block (doubles), canonical orientation: a = offset along the
//      stripe direction ("row"), b = offset across stripes ("column") ----
constexpr int OFF_W1F = 0;      // [b][a][ci=2][o=8]   conv1 forward
constexpr int OFF_B1  = 144;    // [q=4][o=8]          conv1 bias incl. the constant (cos,sin)=(1,0) taps, per column class
constexpr int OFF_W2F = 176;    // [ci=8][a][b][o=8]   conv2 forward
constexpr int OFF_B2  = 752;    // [o=8]
constexpr int OFF_W3F = 760;    // [ci=8][a][b][o=4]   conv3 forward (o padded 3->4)
constexpr int OFF_B3  = 1048;   // [o=4]
constexpr int OFF_W3T = 1052;   // [o=3][a][k=3][ci=8] conv3 transposed (input gradient)
constexpr int OFF_W2T = 1268;   // [o=8][a][b][ci=8]   conv2 transposed
constexpr int OFF_W1T = 1844;   // [o=8][a][b][ci=2]   conv1 transposed
constexpr int PACK_DOUBLES = 1988;

// The activation planes are [channel][column][row]; a channel stride of V (resp. 3V/4) doubles is a multiple of the 32
// shared-memory banks, which makes the 4-channel tensor-core fragments of ph_conv2_mma / ph_conv2T_mma 4-way bank
// conflicted.  PLANE_PAD extra doubles per channel shift consecutive channels by 8 banks: conflict-free.
constexpr int PLANE_PAD = 4;
FT_HD size_t plane_a_doubles(size_t V) { return NH * (V + PLANE_PAD); }            // A: 8 channels x (V + pad)
FT_HD size_t plane_b_doubles(size_t V) { return NH * (3 * V / 4 + PLANE_PAD); }    // B, C: 8 channels x (3V/4 + pad)

constexpr double PI_D = 3.141592653589793;       // == math.pi == np.pi
constexpr double TWO_PI_D = 6.283185307179586;   // == 2*math.pi (exact doubling)

enum Activation { ACT_SILU = 0, ACT_LEAKY = 1, ACT_RELU = 2 };

// ------------------------------------------------------------------------------------------------
// scalar helpers
// ------------------------------------------------------------------------------------------------

// x - floor(x/2pi)*2pi in [0,2pi): what torch.remainder(x, 2pi) returns (fmod, then +2pi if negative).
// One fma against the double 2pi is exact whenever the quotient is right (the remainder of a double by a
// double is representable); the two fix-ups catch a quotient that rounded across an integer.  The
// library fmod() is a bit-serial loop of ~100 instructions and sat on every active site and on every
// bisection iteration.
FT_HD double rem_2pi(double x) {
    const double k = floor(x * 0.15915494309189535);      // 1/(2pi)
    double r = fma(-k, TWO_PI_D, x);
    if (r < 0.0) r += TWO_PI_D;
    if (r >= TWO_PI_D) r -= TWO_PI_D;
    return r;
}
// torch_mod: convention 0 -> [0,2pi) (ipynb/field_transformation.py:17-18); 1 -> remainder(x+pi,2pi)-pi
FT_HD double mod_2pi(double x, int conv) {
    const double sh = conv == 0 ? 0.0 : PI_D;            // (a select, not a branch; x + 0.0 and r - 0.0 are exact)
    return rem_2pi(x + sh) - sh;
}

// hmc_2dU1.py:127-129
FT_HD double regularize1(double f) {
    double g = (f - PI_D) / TWO_PI_D;
    return TWO_PI_D * (g - floor(g) - 0.5);
}
// The same value, bit for bit, without the division routine (k_chain_plain): with C = RN(1 / 2pi), q0 = RN(a C) is a faithful
// quotient, r = a - q0 * 2pi is exact in one fma, and RN(q0 + r C) is the correctly rounded a / 2pi (Markstein's theorem;
// checked against the division on 4e8 random arguments).  Non-finite arguments give NaN either way.
FT_HD double regularize1_fast(double f) {
    const double a = f - PI_D, q0 = a * 0.15915494309189535, r = fma(-q0, TWO_PI_D, a), g = fma(r, 0.15915494309189535, q0);
    return TWO_PI_D * (g - floor(g) - 0.5);
}

// e^x with ~1 ulp error for |x| <= 708 (finite and monotone-saturated beyond): x = (64 k + j) ln2/64 + r with |r| <= ln2/128, e^x = 2^k * 2^(j/64) * e^r.
// 2^(j/64) comes from a 64-entry table (on the device: 512 bytes of shared memory), the power of two goes into its
// exponent bits with an integer add, and e^r - 1 needs a degree-5 Taylor polynomial only (remainder r^6/6! < 3.5e-17).
// 6 fp64 operations after the range reduction instead of the 14 of a degree-13 polynomial on |r| <= ln2/2: the SiLU
// evaluations are a sixth of the trajectory's run time.
// The scalar constants sit in constant memory on the device (a DFMA takes a constant-bank operand directly).
//
// CONTRACT with the kernels: the dynamic shared memory starts with FT_SMEM_PREFIX doubles of scratch -- [0, 64) the
// executors' reduction slots, [48, 56) atan2x2_fast's theta_1..theta_8, [56, 64) transaction barriers, [64, 128) FT_EXP_TABLE (the chain kernels copy it there before
// any phase runs, see fthmc_capi.cu); the engine's arena follows.
constexpr int FT_SMEM_PREFIX = 128;
constexpr int FT_EXP_TAB_OFF = 64;
#define FT_EXP_TABLE { 1.0, 1.0108892860517005, 1.0218971486541166, 1.0330248790212284, 1.0442737824274138, 1.0556451783605572, 1.0671404006768237, 1.0787607977571199, 1.0905077326652577, 1.102382583307841, 1.1143867425958924, 1.1265216186082418, 1.1387886347566916, 1.1511892299529827, 1.1637248587775775, 1.1763969916502812, \
                       1.189207115002721, 1.202156731452703, 1.215247359980469, 1.22848053610687, 1.241857812073484, 1.255380757024691, 1.2690509571917332, 1.2828700160787783, 1.2968395546510096, 1.3109612115247644, 1.3252366431597413, 1.339667524053303, 1.3542555469368927, 1.3690024229745905, 1.383909881963832, 1.3989796725383112, \
                       1.4142135623730951, 1.42961333839197, 1.4451808069770467, 1.460917794180647, 1.4768261459394993, 1.4929077282912648, 1.5091644275934228, 1.5255981507445384, 1.5422108254079407, 1.559004400237837, 1.5759808451078865, 1.593142151342267, 1.6104903319492543, 1.6280274218573478, 1.645755478153965, 1.6636765803267364, \
                       1.681792830507429, 1.7001063537185235, 1.718619298122478, 1.7373338352737062, 1.7562521603732995, 1.7753764925265212, 1.7947090750031072, 1.8142521755003989, 1.8340080864093424, 1.8539791250833855, 1.8741676341103, 1.8945759815869656, 1.9152065613971474, 1.9360617934922943, 1.9571441241754002, 1.978456026387951 }
#define FT_EXP_COEFS { 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5, \
                       92.33248261689366, 6755399441055744.0, -0.04332169877307024 / 4, -1.1926343307941173e-11 / 4 }
#ifdef __CUDACC__
__constant__ double c_exp[8] = FT_EXP_COEFS;
__constant__ double c_exp_tab[64] = FT_EXP_TABLE;
#endif
FT_HD double exp_fast(double x) {
#ifdef __CUDA_ARCH__
    const double* K = c_exp;
    extern __shared__ __align__(16) double fthmc_dyn_smem[];
    const double* TAB = fthmc_dyn_smem + FT_EXP_TAB_OFF;
#else
    const double K[8] = FT_EXP_COEFS;
    static const double TAB[64] = FT_EXP_TABLE;
#endif
    const double nm = fma(x, K[4], K[5]);                                    // rint(64 x / ln2) in the low mantissa bits
    const double n = nm - K[5];
    double r = fma(n, K[6], x);                                              // Cody-Waite: ln2/64 in two pieces
    r = fma(n, K[7], r);
    double w = K[0];                                                         // e^r = 1 + r (1 + r w),  w = 1/2 + r/6 + r^2/24 + r^3/120
#pragma unroll
    for (int i = 1; i < 4; ++i) w = fma(w, r, K[i]);
    const double q = r * fma(r, w, 1.0);                                     // e^r - 1
    // 2^k * 2^(j/64): k goes into the exponent bits of the table entry.  |x| <= 708 keeps the exponent in range; beyond
    // that k saturates (integer min/max, off the critical path) so that the result stays finite and SiLU keeps its
    // limits (0 and z) instead of producing garbage bits.  A NaN propagates.
    int ni;
#ifdef __CUDA_ARCH__
    ni = __double2loint(nm);                                                 // n as a two's complement integer
#else
    ni = (int)(long long)fmin(fmax(n, -2147483647.0), 2147483647.0);
#endif
    const double tj = TAB[ni & 63];
    int k = ni >> 6;
    k = k < -1022 ? -1022 : (k > 1021 ? 1021 : k);
    double sc;
#ifdef __CUDA_ARCH__
    sc = __hiloint2double(__double2hiint(tj) + (k << 20), __double2loint(tj));
#else
    long long bits;
    memcpy(&bits, &tj, sizeof(bits));
    bits += (long long)k << 52;
    memcpy(&sc, &bits, sizeof(sc));
#endif
    return fma(sc, q, sc);
}

// cos for the Wilson action stencil: x = n pi + r, |r| <= pi/2, by a three-term Cody-Waite reduction (pi in 34-bit pieces,
// twice fdlibm's pio2_1 / pio2_2 / pio2_2t), then cos x = (-1)^n cos r with ONE even polynomial (Taylor to r^20: truncation
// < 2e-17 on the interval) -- 16 fp64 operations, every Horner step with a single constant-bank operand, the sign through
// the parity bit of n: no per-lane coefficient selects, no constant loads into registers.  (The first form reduced to
// [-pi/4, pi/4] and blended sine / cosine coefficients per lane: 21 fp64 operations, eight LDC.64 and six FSEL per site; the
// reduction scans are ISSUE bound, not HBM bound.)  Absolute error ~1.5e-16 (the result is not relatively accurate next to
// its zeros, which a sum of cosines does not need).  |x| < 2^19; beyond that, and for non-finite arguments, the library.
// constants: [0] 1/pi, [1] 1.5 * 2^52, [2..4] -pi in three pieces, [5..14] 1/20!, -1/18!, ... 1/4!, -1/2
// (sincos_fast's constants: [0] 2/pi, [1] 1.5 * 2^52, [2..4] -pi/2 in three pieces, [5..10] fdlibm's sine kernel S6..S1)
#define FT_TRIG_CONSTS { 0.63661977236758134308, 6755399441055744.0, \
    -1.57079632673412561417e+00, -6.07710050630396597660e-11, -2.02226624879595063154e-21, \
    1.58969099521155010221e-10, -2.50507602534068634195e-08, 2.75573137070700676789e-06, \
    -1.98412698298579493134e-04, 8.33333333332248946124e-03, -1.66666666666666324348e-01 }
#define FT_COSPI_CONSTS { 0.31830988618379067154, 6755399441055744.0, \
    -2.0 * 1.57079632673412561417e+00, -2.0 * 6.07710050630396597660e-11, -2.0 * 2.02226624879595063154e-21, \
    1.0 / 2432902008176640000.0, -1.0 / 6402373705728000.0, \
    1.0 / 20922789888000.0, -1.0 / 87178291200.0, 1.0 / 479001600.0, -1.0 / 3628800.0, 1.0 / 40320.0, -1.0 / 720.0, 1.0 / 24.0, -0.5 }
// sin_fast: the same reduction, sin x = (-1)^n sin r, sin r = r + r s Q(s) with the odd Taylor series to r^21 (truncation 1.2e-18)
#define FT_SINPI_CONSTS { 1.0 / 51090942171709440000.0, -1.0 / 121645100408832000.0, 1.0 / 355687428096000.0, -1.0 / 1307674368000.0, \
    1.0 / 6227020800.0, -1.0 / 39916800.0, 1.0 / 362880.0, -1.0 / 5040.0, 1.0 / 120.0, -1.0 / 6.0 }
#ifdef __CUDACC__
__constant__ double c_trig[11] = FT_TRIG_CONSTS;
__constant__ double c_cospi[15] = FT_COSPI_CONSTS;
__constant__ double c_sinpi[10] = FT_SINPI_CONSTS;
#endif
FT_HD double cos_core(double x);
FT_HD double cos_fast(double x) { return fabs(x) < 524288.0 ? cos_core(x) : cos(x); }
// the polynomial path alone, |x| < 2^19 (the scans test the sites of a vector together: one branch per vector)
FT_HD double cos_core(double x) {
#ifdef __CUDA_ARCH__
    const double* K = c_cospi;
#else
    const double K[15] = FT_COSPI_CONSTS;
#endif
    const double t = fma(x, K[0], K[1]);                                      // n = rint(x / pi) in the low mantissa bits
    const double n = t - K[1];
    double r = fma(n, K[2], x);
    r = fma(n, K[3], r);
    r = fma(n, K[4], r);
    const double s = r * r;
    double p = K[5];
#pragma unroll
    for (int i = 6; i < 15; ++i) p = fma(p, s, K[i]);
    p = fma(p, s, 1.0);
#ifdef __CUDA_ARCH__
    return __hiloint2double(__double2hiint(p) ^ (__double2loint(t) << 31), __double2loint(p));
#else
    return (((long long)n) & 1) ? -p : p;
#endif
}
#ifndef FT_FAST_SIN
#define FT_FAST_SIN 1
#endif
// sin for the Wilson force (k_force, the plain-HMC leapfrog): 18 fp64 operations; absolute error ~2e-16, relative accuracy
// next to the zeros at multiples of pi as well (the result is r (1 + ...) there).
FT_HD double sin_core(double x);
FT_HD double sin_fast(double x) { return fabs(x) < 524288.0 ? sin_core(x) : sin(x); }
FT_HD double sin_core(double x) {                                              // the polynomial path alone, |x| < 2^19
#ifdef __CUDA_ARCH__
    const double* K = c_cospi; const double* S = c_sinpi;
#else
    const double K[15] = FT_COSPI_CONSTS; const double S[10] = FT_SINPI_CONSTS;
#endif
    const double t = fma(x, K[0], K[1]);
    const double n = t - K[1];
    double r = fma(n, K[2], x);
    r = fma(n, K[3], r);
    r = fma(n, K[4], r);
    const double s = r * r;
    double p = S[0];
#pragma unroll
    for (int i = 1; i < 10; ++i) p = fma(p, s, S[i]);                          // Q(s) = -1/6 + s/120 - ...
    p = fma(r * s, p, r);
#ifdef __CUDA_ARCH__
    return __hiloint2double(__double2hiint(p) ^ (__double2loint(t) << 31), __double2loint(p));
#else
    return (((long long)n) & 1) ? -p : p;
#endif
}
// the sine of the plain-HMC kernels (k_force, k_chain_plain).  The flow kernels keep the library sine in their Wilson-force
// seed (0.4 % of a force evaluation): their register allocation is fragile -- the same seed with sin_fast inlined made k_chain
// 0.9 % slower as a whole, a kernel without the training sweeps 6 % slower (profiles/r2_microopt_ab.txt (16), (17)).
FT_HD double sin_force(double x) { return FT_FAST_SIN ? sin_fast(x) : sin(x); }

// sin(x) and cos(x) together: one three-term Cody-Waite reduction x = n pi/2 + r, fdlibm's sine and cosine kernels on the
// remainder, quadrant swap / signs by selects.  Branch-free for |x| < 2^19 (~24 fp64 operations; ~1 ulp), so that the
// compiler interleaves it with the neighbouring exp / atan chains of a site; the library sincos() beyond.
// constants: FT_TRIG_CONSTS and the cosine kernel C6..C1
#define FT_COS_CONSTS { -1.13596475577881948265e-11, 2.08757232129817482790e-09, -2.75573143513906633035e-07, \
                        2.48015872894767294178e-05, -1.38888888888741095749e-03, 4.16666666666666019037e-02 }
#ifdef __CUDACC__
__constant__ double c_cosk[6] = FT_COS_CONSTS;
#endif
// RANGE_OK: the caller has tested |x| < 2^19 itself.  The per-site phases of the engine test ONCE per site and run the whole
// site's arithmetic in one branch-free block (site_fast / site_slow below): a range test inside this function splits the
// block, and the scheduler does not interleave the sine / cosine chain with the exponential chains around it across the split.
struct TrigFast { static constexpr bool value = true; };
struct TrigChecked { static constexpr bool value = false; };
template <bool RANGE_OK = false>
FT_HD void sincos_fast(double x, double& sn, double& cs) {
    if (!RANGE_OK) { if (!(fabs(x) < 524288.0)) { sn = sin(x); cs = cos(x); return; } }
#ifdef __CUDA_ARCH__
    const double* K = c_trig; const double* C = c_cosk;
#else
    const double K[11] = FT_TRIG_CONSTS; const double C[6] = FT_COS_CONSTS;
#endif
    const double t = fma(x, K[0], K[1]);
    const double n = t - K[1];
    int q;
#ifdef __CUDA_ARCH__
    q = __double2loint(t);
#else
    q = (int)(long long)n;
#endif
    double r = fma(n, K[2], x);
    r = fma(n, K[3], r);
    r = fma(n, K[4], r);
    const double r2 = r * r;
    double ps = K[5], pc = C[0];
#pragma unroll
    for (int i = 1; i < 6; ++i) { ps = fma(ps, r2, K[5 + i]); pc = fma(pc, r2, C[i]); }
    const double s0 = fma(r * r2, ps, r);                                      // sin r
    const double c0 = fma(r2 * r2, pc, fma(-0.5, r2, 1.0));                    // cos r
    const bool odd = q & 1;
    const double sv = odd ? c0 : s0, cv = odd ? s0 : c0;
    sn = (q & 2) ? -sv : sv;
    cs = ((q + 1) & 2) ? -cv : cv;
}

// n / d for a normal d away from over/underflow: hardware reciprocal seed, one cubic Newton step, one residual correction
// (error < 1 ulp).  7 fp64 operations and no branch, where the compiler's division is a call with a slow path.
FT_HD double div_fast(double n, double d) {
#ifdef __CUDA_ARCH__
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double e = fma(-d, y, 1.0);
    y = fma(y, fma(e, e, e), y);
    const double q = n * y;
    return fma(fma(-d, q, n), y, q);
#else
    return n / d;
#endif
}

// 2 atan2(y, x) in [-2pi, 2pi] for (x, y) != (0, 0), branch-free (~30 fp64 operations, ~2 ulp): with mx = max(|x|,|y|),
// mn = min(|x|,|y|), the octant angle atan(mn/mx) = theta_j + atan((mn - mx j/8) / (mx + mn j/8)), theta_j = atan(j/8),
// j = rint(8 mn/mx) from a single-precision estimate: ONE division, its quotient below 1/15, a six-term odd polynomial.
// The mixture transform needs exactly this: mod(2 atan(e^s tan(u/2))) = mod(2 atan2(e^s sin(u/2), cos(u/2))) -- no
// tangent, no pole at u = pi.
#define FT_ATAN_TAB { 0.0, 0.12435499454676144, 0.24497866312686414, 0.35877067027057225, 0.46364760900080615, \
                      0.5585993153435624, 0.6435011087932844, 0.7188299996216245, 0.7853981633974483 }
#define FT_ATAN_COEFS { 1.0 / 13.0, -1.0 / 11.0, 1.0 / 9.0, -1.0 / 7.0, 1.0 / 5.0, -1.0 / 3.0 }
constexpr int FT_ATAN_TAB_OFF = 47;     // theta_1..theta_8 sit in slots [48, 56) of the shared scratch prefix (see FT_SMEM_PREFIX)
#ifdef __CUDACC__
__constant__ double c_atan[6] = FT_ATAN_COEFS;
__constant__ double c_atan_tab[9] = FT_ATAN_TAB;
#endif
FT_HD double atan2x2_fast(double y, double x) {
#ifdef __CUDA_ARCH__
    const double* A = c_atan;
    extern __shared__ __align__(16) double fthmc_dyn_smem[];
    const double* TAB = fthmc_dyn_smem + FT_ATAN_TAB_OFF;
#else
    const double A[6] = FT_ATAN_COEFS;
    static const double TAB[9] = FT_ATAN_TAB;
#endif
    const double ay = fabs(y), ax = fabs(x);
    const bool steep = ay > ax;
    const double mx = steep ? ay : ax, mn = steep ? ax : ay;
    int j;
#ifdef __CUDA_ARCH__
    float rf;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"(__double2float_rn(mx)));
    j = __float2int_rn(8.0f * __double2float_rn(mn) * rf);
#else
    j = (int)lrint(8.0 * mn / mx);
#endif
    j = j < 0 ? 0 : (j > 8 ? 8 : j);
    const double jf = 0.125 * (double)j;
    const double t = div_fast(fma(-mx, jf, mn), fma(mn, jf, mx));
    const double t2 = t * t;
    double p = A[0];
#pragma unroll
    for (int i = 1; i < 6; ++i) p = fma(p, t2, A[i]);
    double a = fma(t * t2, p, t) + (j ? TAB[j] : 0.0);
    if (steep) a = 1.5707963267948966 - a;
    if (x < 0.0) a = PI_D - a;
    a = a + a;
    return y < 0.0 ? -a : a;
}
// torch_mod of a value already within [-2pi, 2pi]
FT_HD double mod_2pi_near(double g, int conv) {
    // (selects, no branch on the convention: a branch splits the branch-free site blocks of the phases that call this)
    const double lo = conv == 0 ? 0.0 : -PI_D, hi = conv == 0 ? TWO_PI_D : PI_D;
    const double r = g < lo ? g + TWO_PI_D : g;
    return r >= hi ? r - TWO_PI_D : r;
}

// 1/d for d >= 1 (one cubic Newton step on the hardware seed: three DFMA); host: plain division
FT_HD double rcp_ge1(double d) {
#ifdef __CUDA_ARCH__
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double e = fma(-d, y, 1.0);             // seed error e ~ 2^-20; y (1 + e + e^2) leaves e^3
    y = fma(y, fma(e, e, e), y);
    return y;
#else
    return 1.0 / d;
#endif
}

// activation (and derivative) with the kind as a compile-time parameter: the element loops below are
// unrolled and must stay branch-free so that several exp chains interleave in one warp
#ifndef FT_SILU_V2
#define FT_SILU_V2 1
#endif
#ifndef FT_CL_PLANES
#define FT_CL_PLANES 1
#endif
#ifndef FT_C3_UNROLL
#define FT_C3_UNROLL 8
#endif
#ifndef FT_CL_SAVED_U
#define FT_CL_SAVED_U 1
#endif
// 1 + e^x for the SiLU passes (issue bound: every fp64 operation counts): exp_fast with the argument reduction against ONE
// constant (the fma is exact; the error n * |C - ln2/64| <= 8.7e-19 n stays below 1 ulp of e^x for |x| < 3.5 and grows to
// ~6 ulp at |x| = 20 -- sigma(z) is then within 2e-9 of 0 or 1), and the "+ 1" folded into the last fma's addend so that it
// leaves the dependent chain: 10 fp64 operations instead of 12.
FT_HD double one_plus_exp(double x) {
#ifdef __CUDA_ARCH__
    const double* K = c_exp;
    extern __shared__ __align__(16) double fthmc_dyn_smem[];
    const double* TAB = fthmc_dyn_smem + FT_EXP_TAB_OFF;
#else
    const double K[8] = FT_EXP_COEFS;
    static const double TAB[64] = FT_EXP_TABLE;
#endif
    const double nm = fma(x, K[4], K[5]);
    const double n = nm - K[5];
    const double r = fma(n, -0.010830424696249145, x);                        // -ln2/64
    double w = K[0];
#pragma unroll
    for (int i = 1; i < 4; ++i) w = fma(w, r, K[i]);
    const double q = r * fma(r, w, 1.0);
    int ni;
#ifdef __CUDA_ARCH__
    ni = __double2loint(nm);
#else
    ni = (int)(long long)fmin(fmax(n, -2147483647.0), 2147483647.0);
#endif
    const double tj = TAB[ni & 63];
    int k = ni >> 6;
    k = k < -1022 ? -1022 : (k > 1021 ? 1021 : k);
    double sc;
#ifdef __CUDA_ARCH__
    sc = __hiloint2double(__double2hiint(tj) + (k << 20), __double2loint(tj));
#else
    long long bits;
    memcpy(&bits, &tj, sizeof(bits));
    bits += (long long)k << 52;
    memcpy(&sc, &bits, sizeof(sc));
#endif
    return fma(sc, q, sc + 1.0);
}
template <int ACT> FT_HD void act_fwd_t(double z, double& h) {
    if (ACT == ACT_SILU) h = z * rcp_ge1(FT_SILU_V2 ? one_plus_exp(-z) : 1.0 + exp_fast(-z));
    else if (ACT == ACT_LEAKY) h = z > 0.0 ? z : 0.01 * z;
    else h = z > 0.0 ? z : 0.0;
}
template <int ACT> FT_HD void act_fwd_der_t(double z, double& h, double& d) {
    if (ACT == ACT_SILU) {
        double sg = rcp_ge1(FT_SILU_V2 ? one_plus_exp(-z) : 1.0 + exp_fast(-z));
        h = z * sg;
        d = fma(sg, fma(-z, sg, z), sg);          // sg * (1 + z * (1 - sg))
    } else if (ACT == ACT_LEAKY) {
        h = z > 0.0 ? z : 0.01 * z;
        d = z > 0.0 ? 1.0 : 0.01;
    } else {
        h = z > 0.0 ? z : 0.0;
        d = z > 0.0 ? 1.0 : 0.0;
    }
}
// in-place activation of N of this thread's own values buf[index(e)], e = 0..N-1; dsave != nullptr: the
// derivative goes to the same index of the global layer block.  Blocks of U elements: all loads, then U
// independent chains, then all stores (a load after a store to the same array would serialise them).
template <int ACT, int N, class IndexFn> FT_HD void act_pass(double* buf, double* dsave, double* hsave, IndexFn index) {
    constexpr int U = 4;      // (blocks of 8, and a two-stage software pipeline of the exp / reciprocal halves, measured no faster)
    static_assert(N % U == 0, "element count must be a multiple of the block");
#pragma unroll 1
    for (int e0 = 0; e0 < N; e0 += U) {
        int idx[U]; double z[U], h[U], d[U];
#pragma unroll
        for (int j = 0; j < U; ++j) { idx[j] = index(e0 + j); z[j] = buf[idx[j]]; }
        if (dsave) {
#pragma unroll
            for (int j = 0; j < U; ++j) act_fwd_der_t<ACT>(z[j], h[j], d[j]);
#pragma unroll
            for (int j = 0; j < U; ++j) { buf[idx[j]] = h[j]; dsave[idx[j]] = d[j]; }
            if (hsave) {
#pragma unroll
                for (int j = 0; j < U; ++j) hsave[idx[j]] = h[j];
            }
        } else {
#pragma unroll
            for (int j = 0; j < U; ++j) act_fwd_t<ACT>(z[j], h[j]);
#pragma unroll
            for (int j = 0; j < U; ++j) buf[idx[j]] = h[j];
        }
    }
}
template <int N, class IndexFn> FT_HD void act_pass_any(int act, double* buf, double* dsave, double* hsave, IndexFn index) {
    if (act == ACT_SILU) act_pass<ACT_SILU, N>(buf, dsave, hsave, index);
    else if (act == ACT_LEAKY) act_pass<ACT_LEAKY, N>(buf, dsave, hsave, index);
    else act_pass<ACT_RELU, N>(buf, dsave, hsave, index);
}

struct alignas(16) dbl2 { double x, y; };
FT_HD dbl2 ld2(const double* p) { return *reinterpret_cast<const dbl2*>(p); }
FT_HD void st2(double* p, double x, double y) { dbl2 v; v.x = x; v.y = y; *reinterpret_cast<dbl2*>(p) = v; }

// the same map from sh = sin(x/2), ch = cos(x/2) through atan2 (see atan2x2_fast): what the hot paths evaluate
FT_HD double mixture_fwd_sc(double sh, double ch, double es0, double es1, int conv) {
    const double g0 = mod_2pi_near(atan2x2_fast(es0 * sh, ch), conv);
    const double g1 = mod_2pi_near(atan2x2_fast(es1 * sh, ch), conv);
    return (g0 + g1) / 2;
}
// mean_k mod(2 atan(e^{s_k} tan(x/2)))   (ipynb/field_transformation.py:249-257), es_k = e^{s_k}
FT_HD double mixture_fwd(double x, double es0, double es1, int conv) {
    double th = tan(x / 2);
    double g0 = mod_2pi(2 * atan(es0 * th), conv);
    double g1 = mod_2pi(2 * atan(es1 * th), conv);
    return (g0 + g1) / 2;
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 counter RNG (throughput mode: momenta and Metropolis uniforms made on the device)
// ------------------------------------------------------------------------------------------------
struct Philox {
    uint32_t k0, k1;
    FT_HD static uint32_t mulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
    FT_HD void gen(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) const {
        uint32_t a = k0, b = k1;
        for (int i = 0; i < 10; ++i) {
            uint32_t h0 = mulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
            uint32_t h1 = mulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
            uint32_t n0 = h1 ^ c1 ^ a, n1 = l1, n2 = h0 ^ c3 ^ b, n3 = l0;
            c0 = n0; c1 = n1; c2 = n2; c3 = n3;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
    }
};
// uniform in (0,1) with 53 random bits
FT_HD double u53(uint32_t hi, uint32_t lo) {
    uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;
    return ((double)v + 0.5) * (1.0 / 9007199254740992.0);
}

// ------------------------------------------------------------------------------------------------
// The engine.  E is the execution policy (device: a CTA; host emulation: one serial thread).
//   E::tid(), E::nt(), E::sync(), E::sum(v), E::maxv(v), E::all(pred) (block/cluster-wide AND, a barrier)
//   E::lsync()                        CTA-local barrier (== sync() without a cluster): for phase boundaries across which
//                                     only this rank's own shared memory is produced and consumed
//   E::kLanes, E::warp(), E::nwarps(), E::lane0(), E::use_mma(), E::mma884(d0, d1, a, b)
//                                     warp-level fp64 tensor-core tile D(8x8) += A(8x4) B(4x8) with the PTX m8n8k4 fragment
//                                     layout.  Device: kLanes == 1, every thread holds its own lane's fragment elements;
//                                     serial emulation: kLanes == 32, one "thread" carries all 32 lanes in arrays
//   E::fine(tasks)                    true when the block has more threads than `tasks` (phases then split tasks finer)
//   E::smem()                         base of the chain's shared-memory arena (address space known to nvcc)
//   E::bar_init(n)                    n transaction barriers (mbarrier) for the bulk copies below
//   E::bulk_load(bar, dst_smem, src_global, ndoubles)   ONE thread: TMA bulk copy (cp.async.bulk) completing on `bar`
//   E::bar_wait(bar, parity)          all threads: wait for the phase of `bar` with that parity
//   E::proxy_fence()                  generic-proxy writes before it become visible to later bulk copies
//
// Cluster mode (E::kCluster, lattices too large for one SM: L = 64 .. 128).  A thread-block cluster of nr =
// E::nranks() CTAs owns one chain.  The link field X and the gradient GR are split by lattice row blocks
// (rank k holds n0 in [k*H, (k+1)*H), H = L0/nr); the per-layer planes (CS, UA, OUT, A, B, C, Pbar) are split
// by canonical column blocks (rank k holds the stripe groups [k*G, (k+1)*G) of the current layer), so every
// convolution tap is local except the two halo columns conv2 / conv2^T need from the neighbouring rank:
// those are pushed through distributed shared memory (E::peer) into a halo buffer before the phase.
// X/GR accesses that fall outside the own row block go through E::peer directly.  E::sync, E::sum and
// E::maxv are cluster-wide in this mode.  With kCluster == false everything below compiles to the
// single-CTA code (rank 0 of 1).
// ------------------------------------------------------------------------------------------------
struct LayerGeom {
    int mu, off;
    int R;     // sites along the stripes
    int Cn;    // sites across the stripes held by this rank (multiple of 4); all of them without a cluster
    int G;     // Cn / 4 stripe groups held by this rank
    int CnG;   // sites across the stripes of the whole lattice
    int c0;    // first canonical column of this rank (before the shift by off)
};

struct EngineParams {
    int L0, L1;
    int nlayers;
    int act, conv;
    double inv_tol;
    int inv_max_iter;
    const double* wpack;   // global: nlayers * PACK_DOUBLES
    const int* lmu;        // global: per layer mu
    const int* loff;       // global: per layer off
    int train;             // 1: the forward sweep also saves the activations h1, h2 and the adjoint accumulates weight gradients
};

// Shared-memory arena (doubles) of one rank; V = sites per rank.  Flow: X GR | CS UA OUT | A B C (8 channels each, padded) | W.
// Plain HMC: X GR | S(V).
FT_HD size_t engine_smem_doubles(int L0, int L1, bool flow = true, int nr = 1) {
    size_t V = (size_t)L0 * L1 / nr, LP = L1 + 1, H = L0 / nr;
    if (!flow) return 2 * H * LP * 2 + V + 4;
    return 2 * H * LP * 2 + V + V / 4 + 3 * (V / 4) + plane_a_doubles(V) + 2 * plane_b_doubles(V) + PACK_DOUBLES + 32 + 4;
}

// Per-layer block of the per-CTA global workspace written by the forward sweep of ft_force and read
// back (bulk copies) by the reverse sweep: act'(z1) [plane A], act'(z2) [plane B], cos/sin of the frozen
// plaquettes [V], (s_1,s_2) of the active sites [2 V/4], pre-update active links [V/4].
// Training mode (weight gradients) appends the activations themselves: h1 [plane A], h2 [plane B].
FT_HD size_t engine_layer_ws_doubles(int L0, int L1, int nr = 1, bool train = false) {
    size_t V = (size_t)L0 * L1 / nr;
    // cluster mode appends the raw active plaquettes [V/4]: the adjoint reads them back instead of re-forming each from four
    // links, most of which it would have to fetch through distributed shared memory
    return (train ? 2 : 1) * (plane_a_doubles(V) + plane_b_doubles(V)) + V + 3 * (V / 4) + (nr > 1 ? V / 4 : 0);
}
// per-chain global workspace (doubles): momenta, x0, y0 (whole lattice), then for every rank nlayers layer blocks
FT_HD size_t engine_ws_doubles(int L0, int L1, int nlayers, int nr = 1, bool train = false) {
    size_t V = (size_t)L0 * L1;
    return 3 * 2 * V + (size_t)nr * nlayers * engine_layer_ws_doubles(L0, L1, nr, train);
}
// weight-gradient accumulators of one warp for one layer, in the layout of the forward half of the packed weights
// (conv1 [b][a][ci][o], S_q[o] in the conv1-bias slots, conv2 [ci][a][b][o], bias2, conv3 [ci][a][b][4], bias3)
constexpr int GRAD_DOUBLES = OFF_W3T;

template <class E>
struct Engine {
    E ex;                 // by value: the engine object itself lives in shared memory on the device, so
                          // that the noinline phases read its members at shared-memory latency
    EngineParams pr;
    static constexpr bool CL = E::kCluster;
    int L0, L1, LP, V, VQ;          // V, VQ: sites per rank (the whole lattice without a cluster)
    int Vg, nr, rk, H;              // whole-lattice volume, ranks in the cluster, own rank, lattice rows per rank
    int sA, sB;                     // channel strides of plane A and of planes B, C (padded, see PLANE_PAD)
    int oX, oGR, oCS, oUA, oOUT, oA, oB, oC, oW, oS, oTab;   // arena offsets (doubles)
    double *wsP, *wsX0, *wsY0;                         // momenta, x0, y0: plane C (single-CTA flow chains) or the global workspace
    double* wsLay;                                     // layer blocks of the adjoint: global per-CTA workspace
    size_t layStride;
    double* gW;                                        // training: this CTA's gradient accumulators [warp][layer][GRAD_DOUBLES]
    int* iters_out;                                    // optional global: bisection iterations per layer
    // transaction barriers of the bulk (TMA) copies.  barcnt[b] counts completed uses: the k-th use of a barrier is
    // waited with parity k & 1.  Thread 0 advances a counter only after a block barrier that follows every thread's wait.
    enum { BAR_W = 0, BAR_D2B, BAR_D2C, BAR_D1, BAR_CS, BAR_SO, BAR_H1, NBAR };
    int barcnt[NBAR];
    // vector-Jacobian mode of the adjoint sweep (fthmc_flow_vjp): an external gradient d/dy of this chain seeds the sweep
    // instead of the Wilson force, and the log-Jacobian terms are weighted by -mw (ft_action = S - sum logJ: mw = 1)
    const double* vjp_seed;
    double mw;

    FT_HD Engine(const E& e, const EngineParams& p, double* ws) : ex(e), pr(p) {
        L0 = p.L0; L1 = p.L1; LP = L1 + 1; Vg = L0 * L1;
        nr = ex.nranks(); rk = ex.rank(); H = L0 / nr; V = Vg / nr; VQ = V / 4;
        sA = V + PLANE_PAD; sB = 3 * VQ + PLANE_PAD;
        int o = 0;
        oX = o;  o += 2 * H * LP;
        oGR = o; o += 2 * H * LP;
        layStride = engine_layer_ws_doubles(L0, L1, nr, p.train != 0);
        gW = nullptr;
        for (int b = 0; b < NBAR; ++b) barcnt[b] = 0;
        wsP = ws; wsX0 = ws + 2 * Vg; wsY0 = ws + 4 * Vg;
        wsLay = ws + 6 * (size_t)Vg + (size_t)rk * p.nlayers * layStride;
        iters_out = nullptr;
        vjp_seed = nullptr; mw = 1.0;
        oCS = oUA = oOUT = oA = oB = oC = oW = oTab = 0;
        if (p.nlayers == 0) { oS = o; return; }        // plain HMC: only a scratch plane
        oCS = o;  o += V;
        oUA = o;  o += VQ;
        oOUT = o; o += 3 * VQ;
        o += (o & 1);                                  // 16-byte alignment of the big planes / weights
        oA = o;   o += NH * sA;
        oB = o;   o += NH * sB;
        oC = o;   o += NH * sB;
        oW = o;   o += PACK_DOUBLES;
        oTab = o; o += 32;                             // per-layer (mu, off) bytes, MAX_LAYERS = 128
        oS = oUA;                                      // Wilson-force scratch plane = UA+OUT (V doubles, contiguous)
        // Single-CTA trajectories keep the momenta and the two saved fields (x0: start of the MD evolution, y0 = F(x0): the
        // field returned on reject) ON CHIP, in plane C (6 V + 32 doubles; free now that act'(z2) is single-buffered): between
        // the load and the store of the field a trajectory touches global memory only for the layer blocks of the adjoint.
        // (Cluster mode keeps its halos and the Pbar transpose in C, the training mode h2: both stay with the global workspace.)
        if (!CL && FT_D2_SINGLE && !p.train) { wsP = ex.smem() + oC; wsX0 = wsP + 2 * Vg; wsY0 = wsP + 4 * Vg; }
    }
    // copy the per-layer mask parameters into shared memory once per kernel (global-latency off the layer loop)
    FT_HD void load_geom_table() {
        unsigned char* tab = reinterpret_cast<unsigned char*>(sm(oTab));
        for (int l = ex.tid(); l < pr.nlayers; l += ex.nt()) { tab[l] = (unsigned char)pr.lmu[l]; tab[128 + l] = (unsigned char)pr.loff[l]; }
        ex.sync();
    }
    FT_HD double* sm(int off) const { return ex.smem() + off; }
    // layer block pieces in the global workspace
    FT_HD double* wsD1(int l) const { return wsLay + (size_t)l * layStride; }
    FT_HD double* wsD2(int l) const { return wsD1(l) + NH * (size_t)sA; }
    FT_HD double* wsCS(int l) const { return wsD2(l) + NH * (size_t)sB; }
    FT_HD double* wsSO(int l) const { return wsCS(l) + V; }
    FT_HD double* wsSV(int l) const { return wsSO(l) + 2 * VQ; }
    FT_HD double* wsH1(int l) const { return wsSV(l) + VQ; }                 // training mode only
    FT_HD double* wsH2(int l) const { return wsH1(l) + NH * (size_t)sA; }

    // ---- geometry ----
    FT_HD LayerGeom geom(int l) const {
        LayerGeom g;
        const unsigned char* tab = reinterpret_cast<const unsigned char*>(sm(oTab));
        g.mu = tab[l]; g.off = tab[128 + l];
        g.R = g.mu == 0 ? L0 : L1;
        g.CnG = g.mu == 0 ? L1 : L0;
        g.Cn = CL ? g.CnG / nr : g.CnG;
        g.G = g.Cn / 4;
        g.c0 = CL ? rk * g.Cn : 0;
        return g;
    }
    // element (mu, n0, n1) of a link-shaped array (X at base oX, GR at base oGR).  Without a cluster: a plain
    // shared-memory address; with one: the owner rank's copy, through distributed shared memory if it is not ours.
    FT_HD double* xat(int base, int mu, int n0, int n1) const {
        if constexpr (!CL) {
            return sm(base) + (mu * L0 + n0) * LP + n1;
        } else {
            const int owner = n0 / H;
            double* p = sm(base) + (mu * H + (n0 - owner * H)) * LP + n1;
            return owner == rk ? p : ex.peer(p, owner);
        }
    }
    // canonical (row r, shifted column c of this rank) -> lattice site
    FT_HD void site(const LayerGeom& g, int r, int c, int& n0, int& n1) const {
        int co = c + g.c0 + g.off; if (co >= g.CnG) co -= g.CnG;
        if (g.mu == 0) { n0 = r; n1 = co; } else { n0 = co; n1 = r; }
    }
    // plaquette angle; order 0: ipynb/field_transformation.py:118-119, order 1: qed_helpers.py:83-86 / hmc_2dU1.py:114-120
    FT_HD double plaq(int base, int n0, int n1, int order) const {
        int n0p = n0 + 1 == L0 ? 0 : n0 + 1, n1p = n1 + 1 == L1 ? 0 : n1 + 1;
        double a = *xat(base, 0, n0, n1), b = *xat(base, 1, n0p, n1), c = *xat(base, 0, n0, n1p), d = *xat(base, 1, n0, n1);
        return order == 0 ? ((a + b) - c) - d : ((a - d) - c) + b;
    }
    // i-th link of this rank (i in [0, 2V)) -> index si in the rank's X/GR arrays and gi in the global (2,L0,L1) layout
    FT_HD void link_map(int i, int& si, int& gi) const {
        const int n1 = i % L1, rest = i / L1;            // rest = mu*H + local row
        si = rest * LP + n1;
        if constexpr (!CL) gi = i;
        else gi = (rest + (rest >= H ? L0 - H : 0) + rk * H) * L1 + n1;
    }
    // i-th site of this rank (i in [0, V)) -> lattice coordinates
    FT_HD void site_map(int i, int& n0, int& n1) const {
        n0 = i / L1; n1 = i - n0 * L1;
        if constexpr (CL) n0 += rk * H;
    }

    // ---- global <-> shared field copies (global layout (2,L0,L1) contiguous) ----
    FT_HD void load_field(int off, const double* g) {
        double* dst = sm(off);
        for (int i = ex.tid(); i < 2 * V; i += ex.nt()) {
            int si, gi; link_map(i, si, gi);
            dst[si] = g[gi];
        }
    }
    FT_HD void store_field(double* g, int off) {
        const double* src = sm(off);
        for (int i = ex.tid(); i < 2 * V; i += ex.nt()) {
            int si, gi; link_map(i, si, gi);
            g[gi] = src[si];
        }
    }
    // stage layer l's weights (one TMA bulk copy on BAR_W): the forward part [0,OFF_W3T) for the
    // forward/reverse sweeps, the transposed part [OFF_W3T,PACK) for the adjoint sweep
    FT_HD void issue_weights(int l, bool transposed) {
        const int lo = transposed ? OFF_W3T : 0, n = transposed ? PACK_DOUBLES - OFF_W3T : OFF_W3T;
        if (l >= 0 && ex.tid() == 0) ex.bulk_load(BAR_W, sm(oW) + lo, pr.wpack + (size_t)l * PACK_DOUBLES + lo, n);
    }
    FT_HD void wait_bar(int b) const { ex.bar_wait(b, barcnt[b] & 1); }
    FT_HD void advance_bar(int b) { if (ex.tid() == 0) barcnt[b] = barcnt[b] + 1; }

    // =============================================================================================
    // plain Wilson action pieces on the resident field
    // =============================================================================================
    // -beta * sum cos P   (order 0: U1GaugeAction, order 1: hmc_2dU1.action)
    FT_PHASE double wilson_action(double beta, int order) {
        double acc = 0.0;
        const int d0 = ex.nt() / L1, d1 = ex.nt() - d0 * L1;
        int n0 = ex.tid() / L1, n1 = ex.tid() - n0 * L1;
        if constexpr (CL) n0 += rk * H;
        for (int i = ex.tid(); i < V; i += ex.nt(), n0 += d0, n1 += d1) {
            if (n1 >= L1) { n1 -= L1; ++n0; }
            acc += cos(plaq(oX, n0, n1, order));
        }
        return -beta * ex.sum(acc);
    }
    // floor(0.1 + sum regularize(P) / 2pi)   hmc_2dU1.py:123-124
    FT_PHASE double topo_floor() {
        double acc = 0.0;
        for (int i = ex.tid(); i < V; i += ex.nt()) { int n0, n1; site_map(i, n0, n1); acc += regularize1(plaq(oX, n0, n1, 1)); }
        return floor(0.1 + ex.sum(acc) / TWO_PI_D);
    }
    // GR = dS/dx of the Wilson action: F0 = beta[sinP(n) - sinP(n-e1)], F1 = beta[sinP(n-e0) - sinP(n)]
    FT_PHASE void wilson_force(double beta, int order) {
        if constexpr (CL) {
            // the sin(P) plane would need a row halo from the neighbouring rank: recompute the two shifted
            // plaquettes instead (three sines per site; this phase is well under 1% of a force evaluation)
            for (int i = ex.tid(); i < V; i += ex.nt()) {
                int n0, n1; site_map(i, n0, n1);
                const int n0m = n0 == 0 ? L0 - 1 : n0 - 1, n1m = n1 == 0 ? L1 - 1 : n1 - 1;
                const double s = sin(plaq(oX, n0, n1, order)), s1 = sin(plaq(oX, n0, n1m, order)), s0 = sin(plaq(oX, n0m, n1, order));
                *xat(oGR, 0, n0, n1) = beta * (s - s1);
                *xat(oGR, 1, n0, n1) = beta * (s0 - s);
            }
            ex.sync();
            return;
        }
        double* S = sm(oS);                              // scratch plane, pitch L1
        // (n0, n1) of the sites a thread visits advance without divisions: i += nt  <=>  n0 += nt / L1, n1 += nt % L1
        const int d0 = ex.nt() / L1, d1 = ex.nt() - d0 * L1, s0 = ex.tid() / L1, s1 = ex.tid() - s0 * L1;
        for (int i = ex.tid(), n0 = s0, n1 = s1; i < V; i += ex.nt(), n0 += d0, n1 += d1) {
            if (n1 >= L1) { n1 -= L1; ++n0; }
            S[i] = sin(plaq(oX, n0, n1, order));
        }
        ex.sync();
        for (int i = ex.tid(), n0 = s0, n1 = s1; i < V; i += ex.nt(), n0 += d0, n1 += d1) {
            if (n1 >= L1) { n1 -= L1; ++n0; }
            int n0m = n0 == 0 ? L0 - 1 : n0 - 1, n1m = n1 == 0 ? L1 - 1 : n1 - 1;
            double s = S[i];
            *xat(oGR, 0, n0, n1) = beta * (s - S[n0 * L1 + n1m]);
            *xat(oGR, 1, n0, n1) = beta * (S[n0m * L1 + n1] - s);
        }
        ex.sync();
    }

    // =============================================================================================
    // coupling-layer phases (canonical stripe geometry)
    // =============================================================================================
    // (cluster mode) scratch behind the conv halos AH / ZH (NH * 2 * R doubles) in arena C: the Pbar transpose of ph_scatter
    FT_HD int oStage() const { return oC + 2 * NH * (L0 > L1 ? L0 : L1); }
    FT_HD static int log2_or_neg(int v) { if (v <= 0 || (v & (v - 1))) return -1; int sh = 0; while ((1 << sh) < v) ++sh; return sh; }
    // plaquette of the canonical site (r, c) of this rank
    FT_HD double plaq_canon(const LayerGeom& g, int r, int c, int order) const {
        int n0, n1; site(g, r, c, n0, n1);
        return plaq(oX, n0, n1, order);
    }

    // cos/sin of the frozen plaquettes and the raw active plaquette; cs_save: global copy of CS
    FT_PHASE void ph_planes(const LayerGeom g, double* cs_save) {
        double* CS = sm(oCS); double* UA = sm(oUA);
        const int T = g.G * g.R, order = pr.conv;
#if FT_CL_PLANES
        if constexpr (CL) {
            // Cluster mode, mu = 0 layers: the stripes run along n0, the links are split by blocks of n0, so all but 1 / nr of
            // the plaquettes of this rank's columns are read through distributed shared memory.  With the lanes of a warp
            // along the stripe (consecutive n0) every lane's 8 bytes come from a different row, often a different owner;
            // remote loads are served a sector at a time.  Two steps instead: the plaquette angles with the lanes ACROSS
            // the stripes (consecutive n1: the rank's Cn columns of a row are contiguous in the owner's memory) into a
            // scratch plane (plane A is free here; pitch R + 1: conflict-free both ways), then cos / sin with the lanes
            // along the stripe as everywhere else.  Same operands, same operations: bit-identical to the direct form.
            if (g.mu == 0 && !fine_tasks()) {
                double* SCR = sm(oA);
                const int Cn = g.Cn, R = g.R, RP = R + 1;
                const int dr = ex.nt() / Cn, dc = ex.nt() - dr * Cn;
                int r = ex.tid() / Cn, c = ex.tid() - r * Cn;
                for (int i = ex.tid(); i < V; i += ex.nt(), r += dr, c += dc) {
                    if (c >= Cn) { c -= Cn; ++r; }
                    if ((c & 3) != 3) SCR[c * RP + r] = plaq_canon(g, r, c, order);
                }
                ex.lsync();
                for (int t = ex.tid(); t < T; t += ex.nt()) {
                    const int gi = t / R, rr = t - gi * R;
                    UA[t] = SCR[(4 * gi) * RP + rr];
#pragma unroll 1
                    for (int k = 0; k < 2; ++k) {
                        const double p = SCR[(4 * gi + 1 + k) * RP + rr];
                        double sp, cp;
                        sincos_fast(p, sp, cp);
                        const int i = (2 * gi + k) * R + rr;
                        CS[i] = cp; CS[V / 2 + i] = sp;
                        if (cs_save) { cs_save[i] = cp; cs_save[V / 2 + i] = sp; }
                    }
                }
                return;
            }
        }
#endif
        if (fine_tasks()) {                                  // one plaquette per task: T active, then 2T frozen
            for (int t = ex.tid(); t < 3 * T; t += ex.nt()) {
                const int kind = t / T, tt = t - kind * T;
                const int gi = tt / g.R, r = tt - gi * g.R;
                const double p = plaq_canon(g, r, 4 * gi + kind, order);
                if (kind == 0) UA[tt] = p;
                else {
                    double sp, cp;
                    sincos_fast(p, sp, cp);
                    const int i = (2 * gi + kind - 1) * g.R + r;
                    CS[i] = cp; CS[V / 2 + i] = sp;
                    if (cs_save) { cs_save[i] = cp; cs_save[V / 2 + i] = sp; }
                }
            }
            return;
        }
        for (int t = ex.tid(); t < T; t += ex.nt()) {
            int gi = t / g.R, r = t - gi * g.R;
            UA[t] = plaq_canon(g, r, 4 * gi, order);
            // the two frozen plaquettes of the row: one range test, then both sine / cosine chains in one branch-free block
            const double p0 = plaq_canon(g, r, 4 * gi + 1, order), p1 = plaq_canon(g, r, 4 * gi + 2, order);
            double sp0, cp0, sp1, cp1;
            if (fabs(p0) < 524288.0 && fabs(p1) < 524288.0) { sincos_fast<true>(p0, sp0, cp0); sincos_fast<true>(p1, sp1, cp1); }
            else { sincos_fast(p0, sp0, cp0); sincos_fast(p1, sp1, cp1); }
            const int i = 2 * gi * g.R + r, j = i + g.R;
            CS[i] = cp0; CS[V / 2 + i] = sp0; CS[j] = cp1; CS[V / 2 + j] = sp1;
            if (cs_save) { cs_save[i] = cp0; cs_save[V / 2 + i] = sp0; cs_save[j] = cp1; cs_save[V / 2 + j] = sp1; }
        }
    }

    // task t of a phase split NS ways per (group, row): lanes stay consecutive in the row r
    template <int NS> FT_HD void task_split(const LayerGeom& g, int t, int& gi, int& h, int& r) const {
        if (NS == 1) { gi = t / g.R; h = 0; r = t - gi * g.R; }
        else { gi = t / (NS * g.R); const int rem = t - gi * NS * g.R; h = rem / g.R; r = rem - h * g.R; }
    }

    // conv1 pre-activations for the 4 columns of group gi at row r.  z[q][o]
    template <int CH> FT_HD void conv1_z(const double* CS, const double* W, const LayerGeom& g, int gi, int r, int h, double z[4][CH]) const {
        const int R = g.R;
        int rr[3] = { r == 0 ? R - 1 : r - 1, r, r + 1 == R ? 0 : r + 1 };
        double in[2][3][2];                                  // [k][a][ci]
#pragma unroll
        for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                in[k][a][0] = CS[(2 * gi + k) * R + rr[a]];
                in[k][a][1] = CS[V / 2 + (2 * gi + k) * R + rr[a]];
            }
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int o = 0; o < CH; o += 2) { const dbl2 bv = ld2(W + OFF_B1 + q * NH + CH * h + o); z[q][o] = bv.x; z[q][o + 1] = bv.y; }
        // output column class q sees frozen column k through kernel column b = k + 2 - q: every weight vector
        // (b, a, ci) is loaded once (128-bit) and feeds the two (q, k) pairs of its kernel column.  Per output the
        // accumulation order is (b ascending, a, ci).
#pragma unroll
        for (int b = 0; b < 3; ++b)
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int ci = 0; ci < 2; ++ci) {
                    double w[CH];
#pragma unroll
                    for (int o = 0; o < CH; o += 2) { const dbl2 wv = ld2(W + OFF_W1F + ((b * 3 + a) * 2 + ci) * NH + CH * h + o); w[o] = wv.x; w[o + 1] = wv.y; }
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        const int q = k + 2 - b;
                        const double v = in[k][a][ci];
#pragma unroll
                        for (int o = 0; o < CH; ++o) z[q][o] = fma(w[o], v, z[q][o]);
                    }
                }
    }

    // h1 = act(conv1) on all columns -> A[o][c][r];  d1_save: act'(z1) to the global layer block
    FT_HD void ph_conv1(const LayerGeom g, double* d1_save, double* h1_save = nullptr) {
        if (fine_tasks()) ph_conv1_t<4>(g, d1_save, h1_save); else ph_conv1_t<8>(g, d1_save, h1_save);
    }
    template <int CH> FT_PHASE void ph_conv1_t(const LayerGeom g, double* d1_save, double* h1_save) {
        const double* CS = sm(oCS); const double* W = sm(oW);
        double* A = sm(oA);
        const int T = g.G * g.R * (NH / CH), R = g.R, act = pr.act;
        for (int t = ex.tid(); t < T; t += ex.nt()) {
            int gi, h, r;
            task_split<NH / CH>(g, t, gi, h, r);
#ifdef FT_PROFILE
            long long tp0 = ex.clock();
#endif
            double z[4][CH];
            conv1_z<CH>(CS, W, g, gi, r, h, z);
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int o = 0; o < CH; ++o) A[(CH * h + o) * sA + (4 * gi + q) * R + r] = z[q][o];
#ifdef FT_PROFILE
            ex.prof_add(PF_C1_MAC, ex.clock() - tp0); tp0 = ex.clock();
#endif
            // activation pass as a partially unrolled loop over this thread's own 32 values (a fully
            // unrolled exp() per element overflows the instruction cache); element e -> channel e/4, column e%4
            const int i0 = CH * h * sA + 4 * gi * R + r, so = sA;
            act_pass_any<4 * CH>(act, A, d1_save, h1_save, [=](int e) { return i0 + (e >> 2) * so + (e & 3) * R; });
#ifdef FT_PROFILE
            ex.prof_add(PF_C1_ACT, ex.clock() - tp0);
#endif
        }
    }

    // ---- cluster halos (distributed shared memory pushes; no-ops without a cluster) ----
    // conv2 of the right neighbour's first group reads our last two columns of h1: AH[ci][j][r] in its arena C
    FT_HD void push_halo_h1(const LayerGeom& g) {
        if constexpr (CL) {
            const double* A = sm(oA);
            double* dst = ex.peer(sm(oC), rk + 1 == nr ? 0 : rk + 1);
            const int R = g.R, n = NH * 2 * R;
            for (int i = ex.tid(); i < n; i += ex.nt()) {
                const int ci = i / (2 * R), rem = i - ci * 2 * R, j = rem / R, r = rem - j * R;
                dst[i] = A[ci * sA + (g.Cn - 2 + j) * R + r];
            }
        }
    }
    // conv2^T of the left neighbour's last group reads columns k = 0, 1 of our first group of zbar2: ZH[o][j][r] in its arena C
    FT_HD void push_halo_zbar2(const LayerGeom& g, int oZ) {
        if constexpr (CL) {
            const double* Z = sm(oZ);
            double* dst = ex.peer(sm(oC), rk == 0 ? nr - 1 : rk - 1);
            const int R = g.R, n = NH * 2 * R;
            for (int i = ex.tid(); i < n; i += ex.nt()) {
                const int o = i / (2 * R), rem = i - o * 2 * R, j = rem / R, r = rem - j * R;
                dst[i] = Z[o * sB + j * R + r];
            }
        }
    }

    // task decomposition of the two big convolutions: t -> (stripe group gi, channel half h, row pair rp).
    // A thread owns rows {2rp, 2rp+1} and 4 of the 8 channels: every weight it loads (128-bit) feeds both
    // rows, every input feeds up to 9 taps x 4 channels, and the two halves of a warp read the same inputs
    // (broadcast).  This keeps the shared-memory pipe under the fp64 pipe.
    // CH = channels per thread: 4 (V/4 tasks per layer) or 2 (V/2 tasks, when the block has more threads than V/4)
    template <int CH> FT_HD void task2(const LayerGeom& g, int t, int& gi, int& h, int& r0) const {
        const int hr = g.R >> 1, per = (NH / CH) * hr;
        gi = t / per;
        const int rem = t - gi * per;
        h = rem / hr;
        r0 = 2 * (rem - h * hr);
    }
    FT_HD bool fine_tasks() const { return ex.fine(VQ); }

    // h2 = act(conv2) on the columns {4g-1,4g,4g+1} -> B[o][3g+k][r];  d2_save: act'(z2) to global
    // tensor-core (DMMA) form of the two big convolutions: stripe lengths that are multiples of 8 (in cluster mode the
    // halo columns of a rank's first / last stripe group come from the halo buffers AH / ZH, channel stride 2R)
    FT_HD bool mma_ok() const { return (L0 & 7) == 0 && (L1 & 7) == 0 && ex.use_mma(); }
    // Winograd F(2,3) along the stripe direction for the two big convolutions (stripes that split into 16-row blocks
    // with at least one block per warp; -DFT_WINOGRAD=0 keeps the direct tensor-core form)
    FT_HD bool wino_ok(const LayerGeom& g) const {
        return FT_WINOGRAD && mma_ok() && (g.R & 15) == 0 && g.G * (g.R >> 4) >= ex.nwarps();
    }
    FT_HD void ph_conv2(const LayerGeom g, double* d2_save, double* h2_save = nullptr) {
        if (wino_ok(g)) {
            if (pr.act == ACT_SILU) ph_conv2_wino<ACT_SILU>(g, d2_save, h2_save);
            else if (pr.act == ACT_LEAKY) ph_conv2_wino<ACT_LEAKY>(g, d2_save, h2_save);
            else ph_conv2_wino<ACT_RELU>(g, d2_save, h2_save);
        } else if (mma_ok()) {
            if (pr.act == ACT_SILU) ph_conv2_mma<ACT_SILU>(g, d2_save, h2_save);
            else if (pr.act == ACT_LEAKY) ph_conv2_mma<ACT_LEAKY>(g, d2_save, h2_save);
            else ph_conv2_mma<ACT_RELU>(g, d2_save, h2_save);
        } else if (fine_tasks()) ph_conv2_t<2>(g, d2_save);      // (the training mode requires the tensor-core path)
        else ph_conv2_t<4>(g, d2_save);
    }

    // conv2 as warp-level GEMM tiles on the fp64 tensor path: D[8 rows of one column][8 output channels] +=
    // A[8 rows][4 input channels of one tap] * B[those 4 channels][8 output channels]; 18 k-chunks = 9 taps x 2 channel
    // halves.  One warp task = 8 rows of one stripe group, all three output columns (three accumulator tiles in flight).
    // The weight fragments stay in registers for the whole phase; A fragments are one 64-bit shared-memory load per
    // lane per DMMA (conflict-free thanks to PLANE_PAD).  Same output layout as ph_conv2_t.
    template <int ACT> FT_PHASE void ph_conv2_mma(const LayerGeom g, double* d2_save, double* h2_save) {
        constexpr int NL = E::kLanes;
        const double* A = sm(oA); const double* W = sm(oW);
        double* B = sm(oB);
        const int R = g.R, Cn = g.Cn, RB = R >> 3;
        double bf[18][NL];                                   // chunk c = tap * 2 + half, tap = a * 3 + b
        FT_LANES(ln, ls) {
            const int j = ln & 3, n = ln >> 2;               // B fragment: row j (input channel 4*half + j), column n (output channel)
#pragma unroll
            for (int c = 0; c < 18; ++c) bf[c][ls] = W[OFF_W2F + ((4 * (c & 1) + j) * 9 + (c >> 1)) * NH + n];
        }
        for (int st = ex.warp(); st < g.G * RB; st += ex.nwarps()) {
            const int gi = st / RB, rb = 8 * (st - gi * RB);
            int cc[5];                                       // columns 4g-2 .. 4g+2 (offsets in doubles)
#pragma unroll
            for (int m = 0; m < 5; ++m) {
                const int c = 4 * gi - 2 + m;
                if (CL && c < 0) cc[m] = (oC - oA) + (c + 2) * R;          // halo AH[ci][c+2][r]
                else cc[m] = (c < 0 ? c + Cn : (c >= Cn ? c - Cn : c)) * R;
            }
            double acc0[3][NL], acc1[3][NL];
            int rowa[3][NL];                                 // A fragment: row i = lane / 4 (site rb + i), column j = lane % 4 (channel)
            int hcj[NL];                                     // cluster mode: channel-stride correction of the halo columns
            FT_LANES(ln, ls) {
                const int i = ln >> 2, j = ln & 3, r = rb + i;
                hcj[ls] = CL ? j * (2 * R - sA) : 0;
                rowa[0][ls] = j * sA + (r == 0 ? R - 1 : r - 1);
                rowa[1][ls] = j * sA + r;
                rowa[2][ls] = j * sA + (r + 1 == R ? 0 : r + 1);
                const double b0 = W[OFF_B2 + 2 * j], b1 = W[OFF_B2 + 2 * j + 1];     // C fragment: channels 2j, 2j+1 of site i
#pragma unroll
                for (int k = 0; k < 3; ++k) { acc0[k][ls] = b0; acc1[k][ls] = b1; }
            }
#ifdef FT_PROFILE
            long long tp0 = ex.clock();
#endif
#pragma unroll
            for (int c = 0; c < 18; ++c) {
                const int tap = c >> 1, a = tap / 3, b = tap - 3 * a, hoff = 4 * (c & 1) * sA;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    double av[NL];
                    const bool halo = CL && gi == 0 && k + b < 2;
                    const int hcorr = halo ? 4 * (c & 1) * (2 * R - sA) : 0;
                    FT_LANES(ln, ls) av[ls] = A[rowa[a][ls] + cc[k + b] + hoff + (halo ? hcj[ls] + hcorr : 0)];
                    ex.mma884(acc0[k], acc1[k], av, bf[c]);
                }
            }
#ifdef FT_PROFILE
            ex.prof_add(PF_C2_MAC, ex.clock() - tp0); tp0 = ex.clock();
#endif
            FT_LANES(ln, ls) {
                const int i = ln >> 2, j = ln & 3;
                const int i0 = 2 * j * sB + 3 * gi * R + rb + i;
                double z[6], h[6], d[6];
#pragma unroll
                for (int k = 0; k < 3; ++k) { z[2 * k] = acc0[k][ls]; z[2 * k + 1] = acc1[k][ls]; }
                if (d2_save) {
#pragma unroll
                    for (int e = 0; e < 6; ++e) act_fwd_der_t<ACT>(z[e], h[e], d[e]);
#pragma unroll
                    for (int e = 0; e < 6; ++e) { const int idx = i0 + (e & 1) * sB + (e >> 1) * R; B[idx] = h[e]; d2_save[idx] = d[e]; }
                    if (h2_save) {
#pragma unroll
                        for (int e = 0; e < 6; ++e) h2_save[i0 + (e & 1) * sB + (e >> 1) * R] = h[e];
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 6; ++e) act_fwd_t<ACT>(z[e], h[e]);
#pragma unroll
                    for (int e = 0; e < 6; ++e) B[i0 + (e & 1) * sB + (e >> 1) * R] = h[e];
                }
            }
#ifdef FT_PROFILE
            ex.prof_add(PF_C2_ACT, ex.clock() - tp0);
#endif
        }
    }
    // conv2 as Winograd F(2,3) along the rows of a stripe on the fp64 tensor path.  A pair of output rows (2i, 2i+1) of
    // one column needs the input rows d0..d3 = 2i-1 .. 2i+2 of the three neighbouring columns:
    //     t = (d0 - d2, d1 + d2, d2 - d1, d1 - d3),   u_b = (g0, (g0+g1+g2)/2, (g0-g1+g2)/2, g2)  (g_a = W[.][a][b][.]),
    //     m_j = sum_{b, ci} t_j[ci][col+b-1] u_b,j[ci][o]   (4 GEMMs of depth 24 instead of 9 taps x 2 rows of depth 8),
    //     out(2i) = m0 + m1 + m2,  out(2i+1) = m1 - m2 - m3.
    // 4 DMMAs per (column tap, channel half) feed TWO rows: 24 DMMAs per 16 sites instead of 36.  One warp task = 8 row
    // pairs of one stripe group, the three output columns (twelve accumulator tiles in flight); the transformed input
    // fragments of a source column are built once (three loads, four DADD) and used by every output column that sees it.
    // The transformed weight fragments are built from the packed forward weights at phase entry and stay in registers.
    template <int ACT> FT_PHASE void ph_conv2_wino(const LayerGeom g, double* d2_save, double* h2_save) {
        constexpr int NL = E::kLanes;
        const double* A = sm(oA); const double* W = sm(oW);
        double* B = sm(oB);
        const int R = g.R, Cn = g.Cn, RB = R >> 4;
        double bf[3][2][4][NL];                              // [column tap b][channel half][j]
        FT_LANES(ln, ls) {
            const int j = ln & 3, n = ln >> 2;               // B fragment: row j (input channel 4*half + j), column n (output channel)
#pragma unroll
            for (int b = 0; b < 3; ++b)
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    const double* wp = W + OFF_W2F + ((4 * hf + j) * 9 + b) * NH + n;
                    const double g0 = wp[0], g1 = wp[3 * NH], g2 = wp[6 * NH], gs = g0 + g2;
                    bf[b][hf][0][ls] = g0; bf[b][hf][1][ls] = 0.5 * (gs + g1); bf[b][hf][2][ls] = 0.5 * (gs - g1); bf[b][hf][3][ls] = g2;
                }
        }
        for (int st = ex.warp(); st < g.G * RB; st += ex.nwarps()) {
            const int gi = st / RB, rb = 16 * (st - gi * RB);
            int cc[5];                                       // columns 4g-2 .. 4g+2 (offsets in doubles)
#pragma unroll
            for (int m = 0; m < 5; ++m) {
                const int c = 4 * gi - 2 + m;
                if (CL && c < 0) cc[m] = (oC - oA) + (c + 2) * R;          // halo AH[ci][c+2][r], channel stride 2R
                else cc[m] = (c < 0 ? c + Cn : (c >= Cn ? c - Cn : c)) * R;
            }
            double acc0[3][4][NL], acc1[3][4][NL];
            int rm[NL], r1[NL], rp[NL];                      // lane (i, j): row pair i, channel j of the half
            int hcj[NL];                                     // cluster mode: channel-stride correction of the halo columns
            FT_LANES(ln, ls) {
                const int i = ln >> 2, j = ln & 3, r0 = rb + 2 * i;
                hcj[ls] = CL ? j * (2 * R - sA) : 0;
                rm[ls] = j * sA + (r0 == 0 ? R - 1 : r0 - 1);
                r1[ls] = j * sA + r0;
                rp[ls] = j * sA + (r0 + 2 == R ? 0 : r0 + 2);
                const double b0 = W[OFF_B2 + 2 * j], b1 = W[OFF_B2 + 2 * j + 1];     // the bias rides on m1 (in both outputs with +1)
#pragma unroll
                for (int k = 0; k < 3; ++k)
#pragma unroll
                    for (int m = 0; m < 4; ++m) { acc0[k][m][ls] = m == 1 ? b0 : 0.0; acc1[k][m][ls] = m == 1 ? b1 : 0.0; }
            }
#ifdef FT_PROFILE
            long long tp0 = ex.clock();
#endif
#pragma unroll
            for (int c = 0; c < 5; ++c)
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    double t[4][NL];
                    const bool halo = CL && gi == 0 && c < 2;
                    FT_LANES(ln, ls) {
                        const double* pc = A + cc[c] + 4 * hf * sA + (halo ? hcj[ls] + 4 * hf * (2 * R - sA) : 0);
                        const double d0 = pc[rm[ls]], d3 = pc[rp[ls]];
                        const dbl2 d12 = ld2(pc + r1[ls]);
                        t[0][ls] = d0 - d12.y; t[1][ls] = d12.x + d12.y; t[2][ls] = d12.y - d12.x; t[3][ls] = d12.x - d3;
                    }
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int b = c - k;
                        if (b >= 0 && b < 3) {
#pragma unroll
                            for (int m = 0; m < 4; ++m) ex.mma884(acc0[k][m], acc1[k][m], t[m], bf[b][hf][m]);
                        }
                    }
                }
#ifdef FT_PROFILE
            ex.prof_add(PF_C2_MAC, ex.clock() - tp0); tp0 = ex.clock();
#endif
            FT_LANES(ln, ls) {
                const int i = ln >> 2, j = ln & 3;
                const int i0 = 2 * j * sB + 3 * gi * R + rb + 2 * i;
                double z[12], h[12], d[12];                  // element e = (k * 2 + channel parity) * 2 + row parity
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    z[4 * k + 0] = (acc0[k][0][ls] + acc0[k][1][ls]) + acc0[k][2][ls];
                    z[4 * k + 1] = (acc0[k][1][ls] - acc0[k][2][ls]) - acc0[k][3][ls];
                    z[4 * k + 2] = (acc1[k][0][ls] + acc1[k][1][ls]) + acc1[k][2][ls];
                    z[4 * k + 3] = (acc1[k][1][ls] - acc1[k][2][ls]) - acc1[k][3][ls];
                }
                if (d2_save) {
#pragma unroll
                    for (int e = 0; e < 12; ++e) act_fwd_der_t<ACT>(z[e], h[e], d[e]);
#pragma unroll
                    for (int e = 0; e < 12; e += 2) {
                        const int idx = i0 + ((e >> 1) & 1) * sB + (e >> 2) * R;
                        st2(B + idx, h[e], h[e + 1]); st2(d2_save + idx, d[e], d[e + 1]);
                    }
                    if (h2_save) {
#pragma unroll
                        for (int e = 0; e < 12; e += 2) st2(h2_save + i0 + ((e >> 1) & 1) * sB + (e >> 2) * R, h[e], h[e + 1]);
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 12; ++e) act_fwd_t<ACT>(z[e], h[e]);
#pragma unroll
                    for (int e = 0; e < 12; e += 2) st2(B + i0 + ((e >> 1) & 1) * sB + (e >> 2) * R, h[e], h[e + 1]);
                }
            }
#ifdef FT_PROFILE
            ex.prof_add(PF_C2_ACT, ex.clock() - tp0);
#endif
        }
    }
    template <int CH> FT_PHASE void ph_conv2_t(const LayerGeom g, double* d2_save) {
        const double* A = sm(oA); const double* W = sm(oW);
        double* B = sm(oB);
        const int T = g.G * g.R * (4 / CH), R = g.R, Cn = g.Cn, act = pr.act;
        for (int t = ex.tid(); t < T; t += ex.nt()) {
            int gi, h, r0;
            task2<CH>(g, t, gi, h, r0);
            const int rm = r0 == 0 ? R - 1 : r0 - 1, rp = r0 + 2 == R ? 0 : r0 + 2;
            // input columns 4g-2 .. 4g+2: offset of channel 0 from A, and the channel stride.  In cluster mode the
            // two columns left of the rank's first group sit in the halo buffer AH[ci][j][r] (arena C)
            int cc[5], cst[5];
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                int c = 4 * gi - 2 + j;
                if (CL && c < 0) { cc[j] = (oC - oA) + (c + 2) * R; cst[j] = 2 * R; }
                else { cc[j] = (c < 0 ? c + Cn : (c >= Cn ? c - Cn : c)) * R; cst[j] = sA; }
            }
            double acc[2][3][CH];
#pragma unroll
            for (int o = 0; o < CH; ++o) {
                const double bv = W[OFF_B2 + CH * h + o];
#pragma unroll
                for (int dr = 0; dr < 2; ++dr)
#pragma unroll
                    for (int k = 0; k < 3; ++k) acc[dr][k][o] = bv;
            }
#ifdef FT_PROFILE
            long long tp0 = ex.clock();
#endif
#pragma unroll 1
            for (int ci = 0; ci < NH; ++ci) {
                double in[4][5];
                const double* Ap = A + ci * sA;
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    const double* col = CL ? A + cc[j] + ci * cst[j] : Ap + cc[j];
                    in[0][j] = col[rm];
                    const dbl2 m = ld2(col + r0);
                    in[1][j] = m.x; in[2][j] = m.y;
                    in[3][j] = col[rp];
                }
                const double* wc = W + OFF_W2F + ci * 9 * NH + CH * h;
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) {
                        double w[CH];
#pragma unroll
                        for (int o = 0; o < CH; o += 2) { const dbl2 wv = ld2(wc + (a * 3 + b) * NH + o); w[o] = wv.x; w[o + 1] = wv.y; }
#pragma unroll
                        for (int dr = 0; dr < 2; ++dr)
#pragma unroll
                            for (int k = 0; k < 3; ++k)
#pragma unroll
                                for (int o = 0; o < CH; ++o) acc[dr][k][o] = fma(w[o], in[dr + a][k + b], acc[dr][k][o]);
                    }
            }
#ifdef FT_PROFILE
            ex.prof_add(PF_C2_MAC, ex.clock() - tp0); tp0 = ex.clock();
#endif
            const int cs = sB, i0 = CH * h * sB + 3 * gi * R + r0;
#pragma unroll
            for (int o = 0; o < CH; ++o)
#pragma unroll
                for (int k = 0; k < 3; ++k) st2(B + i0 + o * cs + k * R, acc[0][k][o], acc[1][k][o]);
            // element e -> (channel e/6, column (e/2)%3, row e%2)
            act_pass_any<6 * CH>(act, B, d2_save, nullptr, [=](int e) { return i0 + (e / 6) * cs + ((e >> 1) % 3) * R + (e & 1); });
#ifdef FT_PROFILE
            ex.prof_add(PF_C2_ACT, ex.clock() - tp0);
#endif
        }
    }

    // conv3 for the two active sites (gi, r), (gi, r+1), r even, of one thread -> OUT[o][t], t = gi R + r.  The phase is
    // bound by shared-memory wavefronts (with one site per thread: three loads per three DFMA, the weight loads warp-uniform): two
    // sites share every weight load and the four input rows r-1 .. r+2 of a column (17 wavefronts per (channel, column)
    // for two sites instead of 30).  Per site one accumulator per (output, kernel row): nine independent chains, summed (bias + row 0) + row 1) + row 2.
    FT_HD void conv3_pair(const double* B, const double* W, double* OUT, const LayerGeom& g, int gi, int r) const {
        const int R = g.R, T = g.G * R;
        const int rm = r == 0 ? R - 1 : r - 1, rp = r + 2 == R ? 0 : r + 2;
        double o0[3][3], o1[3][3];                           // [kernel row a][output]: site r, site r+1
#pragma unroll
        for (int a = 0; a < 3; ++a) { o0[a][0] = o0[a][1] = o0[a][2] = 0.0; o1[a][0] = o1[a][1] = o1[a][2] = 0.0; }
        constexpr int C3U = FT_C3_UNROLL;
#pragma unroll C3U
        for (int ci = 0; ci < NH; ++ci) {
            const double* Bp = B + ci * sB + 3 * gi * R;
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                const dbl2 v12 = ld2(Bp + b * R + r);
                const double v[4] = { Bp[b * R + rm], v12.x, v12.y, Bp[b * R + rp] };     // rows r-1, r, r+1, r+2
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const dbl2 w01 = ld2(W + OFF_W3F + ((ci * 3 + a) * 3 + b) * 4);
                    const double w2 = W[OFF_W3F + ((ci * 3 + a) * 3 + b) * 4 + 2];
                    o0[a][0] = fma(w01.x, v[a], o0[a][0]); o0[a][1] = fma(w01.y, v[a], o0[a][1]); o0[a][2] = fma(w2, v[a], o0[a][2]);
                    o1[a][0] = fma(w01.x, v[a + 1], o1[a][0]); o1[a][1] = fma(w01.y, v[a + 1], o1[a][1]); o1[a][2] = fma(w2, v[a + 1], o1[a][2]);
                }
            }
        }
        const int t = gi * R + r;
#pragma unroll
        for (int o = 0; o < 3; ++o)
            st2(OUT + o * T + t, ((W[OFF_B3 + o] + o0[0][o]) + o0[1][o]) + o0[2][o], ((W[OFF_B3 + o] + o1[0][o]) + o1[1][o]) + o1[2][o]);
    }
    // conv3 of every active site of this rank into OUT (s_1 | s_2 | t planes of T doubles), then a block barrier
    FT_HD void conv3_all(const LayerGeom& g) {
        const double* B = sm(oB); const double* W = sm(oW); double* OUT = sm(oOUT);
        const int T = g.G * g.R, hR = g.R >> 1;
        for (int t2 = ex.tid(); t2 < (T >> 1); t2 += ex.nt()) {
            const int gi = t2 / hR;
            conv3_pair(B, W, OUT, g, gi, 2 * (t2 - gi * hR));
        }
        ex.lsync();
    }

    // forward transform of the active plaquettes + link update; returns this thread's logJ partial.
    // sv/so != nullptr: the pre-update active links and (s_1,s_2) go to the global layer block.
    FT_PHASE double ph_conv3_forward(const LayerGeom g, bool want_logJ, double* sv, double* so) {
        const double* UA = sm(oUA); const double* OUT = sm(oOUT);
        const int T = g.G * g.R, R = g.R, conv = pr.conv;
        double lj = 0.0;
#ifdef FT_PROFILE
        long long tc0 = ex.clock();
#endif
        conv3_all(g);
#ifdef FT_PROFILE
        ex.prof_add(PF_C3_CONV, ex.clock() - tc0);
#endif
        // one site per thread: everything below is branch-free (exp_fast, sincos_fast, atan2x2_fast), so the independent
        // chains of a site -- two exponentials and the half-angle sine/cosine, then the two arctangents -- interleave
        for (int t = ex.tid(); t < T; t += ex.nt()) {
          const double u = UA[t];
          auto site_body = [&](auto trig) {
            int gi = t / R, r = t - gi * R;
            const double out[NOUT] = { OUT[t], OUT[T + t], OUT[2 * T + t] };
            int n0, n1; site(g, r, 4 * gi, n0, n1);
            double* xl = xat(oX, g.mu, n0, n1);
            const double xo = *xl;                            // (cluster mode: possibly a remote load -- in flight under the arithmetic)
            const double es0 = exp_fast(out[0]), es1 = exp_fast(out[1]);
            double sh, ch;
            sincos_fast<decltype(trig)::value>(0.5 * u, sh, ch);
            const double fx1 = mixture_fwd_sc(sh, ch, es0, es1, conv);
            const double newp = mod_2pi(fx1 + out[2], conv);
            const double delta = newp - u;
            if (sv) {
                sv[t] = xo; so[t] = out[0]; so[T + t] = out[1];
                if constexpr (CL) { if (nr > 1) sv[T + t] = u; }   // (cluster mode) the active plaquette, for ph_outgrad_saved
            }
            *xl = mod_2pi((g.mu == 0 ? delta : -delta) + xo, conv);
            if (want_logJ) {
                const double c2 = ch * ch, s2 = sh * sh;
                double l0 = -log(exp_fast(-out[0]) * c2 + es0 * s2);
                double l1 = -log(exp_fast(-out[1]) * c2 + es1 * s2);
                double m = l0 > l1 ? l0 : l1;
                lj += (m + log(exp(l0 - m) + exp(l1 - m))) - 0.6931471805599453;
            }
          };
          if (fabs(u) < 1048576.0) site_body(TrigFast{}); else site_body(TrigChecked{});   // (|u / 2| < 2^19: one test per site)
        }
        return lj;
    }

    // one coupling layer forward on the resident field; returns logJ (valid on all threads) if asked.
    // save: also write the layer block the reverse sweep of ft_force needs.
    // TRAIN (compile time): also save the activations h1, h2 for the weight-gradient phases
    template <bool TRAIN = false> FT_HD double layer_forward(int l, bool want_logJ, bool save) {
        LayerGeom g = geom(l);
        double lj = 0.0, tot = 0.0;
        FT_T(PF_PLANES, issue_weights(l, false);              // lands while the plaquette planes are computed
             ph_planes(g, save ? wsCS(l) : nullptr);
             wait_bar(BAR_W);
             ex.lsync();                                      // CS / UA are produced and consumed by this rank only
             advance_bar(BAR_W));
        const bool tr = TRAIN && save;
        FT_T(PF_CONV1, ph_conv1(g, save ? wsD1(l) : nullptr, tr ? wsH1(l) : nullptr); ex.lsync());
        if constexpr (CL) { push_halo_h1(g); ex.sync(); }     // the neighbour must see the halo: cluster-wide
        FT_T(PF_CONV2, ph_conv2(g, save ? wsD2(l) : nullptr, tr ? wsH2(l) : nullptr); ex.lsync());
        FT_T(PF_CONV3F, lj = ph_conv3_forward(g, want_logJ, save ? wsSV(l) : nullptr, save ? wsSO(l) : nullptr);
             tot = want_logJ ? ex.sum(lj) : 0.0;
             ex.sync());
        return tot;
    }

    // ---- reverse (ipynb/field_transformation.py:168-174, 319-338, 263-285) ----
    // per-chain bisection: the reference path is single-chain, so its tensor-global stop test is a
    // per-chain stop test here.  The "reached floating point precision" exit (:276-279) can never
    // fire within max_iter=1000 because non-active sites (y=0,f=0) keep halving towards 0 without
    // reaching it, so only the tolerance and max_iter exits exist.
    FT_PHASE double ph_conv3_reverse(const LayerGeom g, bool want_logJ, int* iters) {
        const double* UA = sm(oUA);
        double* OUT = sm(oOUT); double* A = sm(oA);
        const int T = g.G * g.R, R = g.R, conv = pr.conv;
        double* Y = A; double* ES0 = A + T; double* ES1 = A + 2 * T; double* LO = A + 3 * T; double* HI = A + 4 * T;
        double* MID = A + 5 * T;
        const double lo0 = conv == 0 ? 0.0 : -PI_D, hi0 = conv == 0 ? TWO_PI_D : PI_D;
        conv3_all(g);
        for (int t = ex.tid(); t < T; t += ex.nt()) {
            ES0[t] = exp_fast(OUT[t]); ES1[t] = exp_fast(OUT[T + t]);
            Y[t] = mod_2pi(UA[t] - OUT[2 * T + t], conv);
            LO[t] = lo0; HI[t] = hi0;
        }
        ex.lsync();    // (A was the conv2 input; all conv3 reads of B are unaffected)
#if FT_BISECT_REPLAY
        // The reference's bisection costs ~23 evaluations of the mixture map f per site and layer.  f is smooth and strictly
        // increasing, so its decisions can be REPLAYED instead of evaluated: a safeguarded Newton iteration first finds the
        // root xs (f(xs) = y up to a residual rho <= 2e-13, ~5 evaluations), then every bisection step forms the same
        // midpoint with the same arithmetic and decides from d = mid - xs:
        //     y - f(mid) = -(f'(xs) d + R),   |R| <= K^3 d^2 / 4 + rho      (|f''| <= K^3 / 2,  K = max_k max(e^s_k, e^-s_k)),
        //     |y - f(mid)| >= mmin |d| - rho,  mmin = mean_k min(e^s_k, e^-s_k) <= f'.
        // When these bounds fix both the sign of y - f(mid) and the side of |y - f(mid)| relative to the tolerance (with a
        // 2e-13 guard for the rounding of an evaluated f), the step needs no evaluation; otherwise -- d within ~1e-12 of
        // zero, the error within ~1e-12 of the tolerance, or a Newton iteration that did not converge -- the step evaluates
        // f(mid) exactly as before.  Midpoints, decisions and the iteration count are those of the plain loop.
        double* XS = A + 6 * T; double* MS = A + 7 * T; double* QC = A + 8 * T; double* RH = A + 9 * T; double* MM = A + 10 * T;
        for (int t = ex.tid(); t < T; t += ex.nt()) {
            const double y = Y[t], es0 = ES0[t], es1 = ES1[t];
            double a = lo0, b = hi0, x = y, fp = 1.0, rho = 1e300;
            if (!(x > a && x < b)) x = 0.5 * (a + b);
#pragma unroll 1
            for (int k = 0; k < 12; ++k) {
                // f through the half-angle sine / cosine (mixture_fwd_sc), f' = mean_k e^s_k / (cos^2 + e^2s_k sin^2) from the same pair
                double sh, ch;
                sincos_fast<true>(0.5 * x, sh, ch);                   // (x lies inside the bisection interval: always in range)
                const double fx = mixture_fwd_sc(sh, ch, es0, es1, conv), r = fx - y;
                const double c2 = ch * ch, s2 = sh * sh;
                const double n0 = fma(es0 * es0, s2, c2), n1 = fma(es1 * es1, s2, c2);
                fp = 0.5 * div_fast(es0 * n1 + es1 * n0, n0 * n1);
                if (fabs(r) <= 2e-13) { rho = fabs(r); break; }
                if (r > 0.0) b = x; else a = x;
                double xn = x - div_fast(r, fp);
                if (!(xn > a && xn < b)) xn = 0.5 * (a + b);
                x = xn;
            }
            const double i0 = 1.0 / es0, i1 = 1.0 / es1;
            const double k0 = es0 > i0 ? es0 : i0, k1 = es1 > i1 ? es1 : i1, K = k0 > k1 ? k0 : k1;
            XS[t] = x; MS[t] = fp; RH[t] = rho + 2e-13;
            QC[t] = 0.25 * K * K * K;
            MM[t] = 0.5 * ((es0 < i0 ? es0 : i0) + (es1 < i1 ? es1 : i1));
        }
#endif
        // one bisection step of site t on [lo, hi]: the midpoint, whether y > f(mid) (the root lies right of it), and
        // whether |y - f(mid)| < tol -- from the bounds when they decide, from an evaluation of f otherwise
        auto site_step = [&](int t, double lo, double hi, double& mid, bool& right) -> bool {
            mid = (lo + hi) / 2;
#if FT_BISECT_REPLAY
            const double d = mid - XS[t], ad = fabs(d), rh = RH[t];
            const double lin = MS[t] * ad, q = QC[t] * ad * ad + rh, glob = MM[t] * ad - rh;
            const double lower = lin - q > glob ? lin - q : glob, upper = lin + q;
            if (lower > 0.0 && (lower >= pr.inv_tol || upper < pr.inv_tol)) {
                right = d < 0.0;                                     // y > f(mid)  <=>  mid left of the root
                return upper < pr.inv_tol;
            }
#endif
            const double f = mixture_fwd(mid, ES0[t], ES1[t], conv), y = Y[t];
            right = y > f;
            return fabs(y - f) < pr.inv_tol;
        };
        int it = 0;
#if FT_BISECT_REPLAY
        // The tensor-wide stop test cannot succeed before every site has been within tolerance once: each site first
        // replays on its own, without barriers, up to its first converged step a_t (not applied); one max-reduction gives
        // A = max_t a_t, the sites catch up to step A, and the barrier loop below runs from there (one to three steps).
        {
            double* AT = A + 11 * T;
            double amax = 0.0;
            for (int t = ex.tid(); t < T; t += ex.nt()) {
                double lo = LO[t], hi = HI[t], mid; bool right;
                int i = 0;
                for (; i < pr.inv_max_iter - 1; ++i) {
                    if (site_step(t, lo, hi, mid, right)) break;
                    if (right) lo = mid; else hi = mid;
                }
                LO[t] = lo; HI[t] = hi; AT[t] = (double)i;
                amax = amax > (double)i ? amax : (double)i;
            }
            it = (int)ex.maxv(amax);
            for (int t = ex.tid(); t < T; t += ex.nt()) {
                double lo = LO[t], hi = HI[t], mid; bool right;
                for (int i = (int)AT[t]; i < it; ++i) {
                    site_step(t, lo, hi, mid, right);
                    if (right) lo = mid; else hi = mid;
                }
                LO[t] = lo; HI[t] = hi;
            }
        }
#endif
        for (; it < pr.inv_max_iter; ++it) {
            bool conv_all = true;
            for (int t = ex.tid(); t < T; t += ex.nt()) {
                double mid; bool right;
                const bool c = site_step(t, LO[t], HI[t], mid, right);
                conv_all = conv_all && c;
                MID[t] = mid;
                if (right) LO[t] = mid; else HI[t] = mid;
            }
            // stop test max_t err_t < tol == AND_t (err_t < tol): one hardware barrier-reduction instead of a shuffle tree,
            // a shared-memory exchange and two barriers per iteration
            if (ex.all(conv_all)) { ++it; break; }
        }
        if (iters) *iters = it;
        double lj = 0.0;
        for (int t = ex.tid(); t < T; t += ex.nt()) {
            int gi = t / R, r = t - gi * R;
            double x1 = MID[t];
            double delta = x1 - UA[t];
            int n0, n1; site(g, r, 4 * gi, n0, n1);
            double* xl = xat(oX, g.mu, n0, n1);
            *xl = mod_2pi((g.mu == 0 ? delta : -delta) + *xl, conv);
            if (want_logJ) {
                double c = cos(x1 / 2), s = sin(x1 / 2);
                double l0 = -log(exp(-OUT[t]) * (c * c) + ES0[t] * (s * s));
                double l1 = -log(exp(-OUT[T + t]) * (c * c) + ES1[t] * (s * s));
                double m = l0 > l1 ? l0 : l1;
                lj -= (m + log(exp(l0 - m) + exp(l1 - m))) - 0.6931471805599453;
            }
        }
        return lj;
    }

    FT_HD double layer_reverse(int l, bool want_logJ) {
        LayerGeom g = geom(l);
        FT_T(PF_PLANES, issue_weights(l, false);
             ph_planes(g, nullptr);
             wait_bar(BAR_W);
             ex.lsync();                                      // CS / UA are produced and consumed by this rank only
             advance_bar(BAR_W));
        FT_T(PF_CONV1, ph_conv1(g, nullptr); ex.lsync());
        if constexpr (CL) { push_halo_h1(g); ex.sync(); }
        FT_T(PF_CONV2, ph_conv2(g, nullptr); ex.lsync());
        int iters = 0;
        double lj = 0.0, tot = 0.0;
        FT_T(PF_CONV3R, lj = ph_conv3_reverse(g, want_logJ, &iters);
             if (iters_out && ex.tid() == 0) iters_out[l] = iters;
             tot = want_logJ ? ex.sum(lj) : 0.0;
             ex.sync());
        return tot;
    }

    // =============================================================================================
    // adjoint of one layer: GR holds d/dy on entry, d/dx on exit; X holds y on entry, x on exit.
    // Nothing of the forward CNN is recomputed: act'(z1), act'(z2), cos/sin of the frozen plaquettes
    // and (s_1,s_2) come back from the layer block written by the forward sweep.
    // =============================================================================================
    // put the pre-update active links back, then the adjoint of the mixture transform and of -logJ at
    // the active sites:  OUT <- (s1bar, s2bar, tbar),  UA <- Pbar(active)
    // On entry OUT holds (s_1, s_2, pre-update active links) of this layer, prefetched from the layer block; every
    // task reads its own three entries before overwriting them.
    FT_PHASE void ph_outgrad(const LayerGeom g) { outgrad_impl<false>(g, nullptr); }
    // cluster mode: the active plaquette comes from the layer block (ua: V/4 doubles behind the saved links) -- with the
    // pre-update link restored it is bit for bit the value the forward sweep formed
    FT_PHASE void ph_outgrad_saved(const LayerGeom g, const double* ua) { outgrad_impl<true>(g, ua); }
    template <bool SAVED> FT_HD void outgrad_impl(const LayerGeom g, const double* ua) {
        double* OUT = sm(oOUT); double* UA = sm(oUA);
        const int T = g.G * g.R, R = g.R, order = pr.conv;
        const double mwl = mw;
        wait_bar(BAR_SO);
        // the active plaquette only involves its own active link, so restore + plaquette fuse per task
        for (int t = ex.tid(); t < T; t += ex.nt()) {
            int gi = t / R, r = t - gi * R, n0, n1;
            site(g, r, 4 * gi, n0, n1);
            *xat(oX, g.mu, n0, n1) = OUT[2 * T + t];
            double u;
            if constexpr (SAVED) u = ua[t]; else u = plaq(oX, n0, n1, order);
          auto site_body = [&](auto trig) {                   // (one branch-free block per site, see sincos_fast)
            double gl = *xat(oGR, g.mu, n0, n1);
            double db = g.mu == 0 ? gl : -gl;                 // delta-bar
            double s0 = OUT[t], s1 = OUT[T + t];
            double c, s;
            sincos_fast<decltype(trig)::value>(0.5 * u, s, c);
            const double su = 2.0 * s * c;                                                  // sin u
            double c2 = c * c, s2 = s * s;
            double ep0 = exp_fast(s0), em0 = exp_fast(-s0), ep1 = exp_fast(s1), em1 = exp_fast(-s1);
            double e0 = div_fast(1.0, em0 * c2 + ep0 * s2), e1 = div_fast(1.0, em1 * c2 + ep1 * s2);   // e^{l_k}
            const double ise = div_fast(1.0, e0 + e1);
            double sg0 = e0 * ise, sg1 = e1 * ise;                                           // softmax_k l_k
            // the logJ terms carry the weight w = -mw (ft_action = S - sum logJ: w = -1, mw = 1; a product with 1.0 is exact)
            double ub = -db + db * (0.5 * (e0 + e1))
                      + mwl * (sg0 * (0.5 * (ep0 - em0)) * su * e0 + sg1 * (0.5 * (ep1 - em1)) * su * e1);
            double sb0 = db * su * e0 * 0.5 - mwl * (sg0 * (em0 * c2 - ep0 * s2) * e0);
            double sb1 = db * su * e1 * 0.5 - mwl * (sg1 * (em1 * c2 - ep1 * s2) * e1);
            OUT[t] = sb0; OUT[T + t] = sb1; OUT[2 * T + t] = db;
            UA[t] = ub;
          };
          if (fabs(u) < 1048576.0) site_body(TrigFast{}); else site_body(TrigChecked{});
        }
    }

    // zbar2 = conv3^T(OUT) * act'(z2)  (in place in C, which holds act'(z2))
    FT_HD void ph_conv3T(const LayerGeom g, int oZ) {
#if FT_CONV3T_PAIR
        if (!fine_tasks() && (g.R & 1) == 0) { ph_conv3T_pair(g, oZ); return; }
#endif
        if (fine_tasks()) ph_conv3T_t<4>(g, oZ); else ph_conv3T_t<8>(g, oZ);
    }
    // two rows (r, r+1), r even, per thread: the 108 warp-uniform weight vectors of a task -- the bulk of the phase's
    // shared-memory wavefronts -- are loaded once for both rows (see ph_conv1T_pair)
    FT_PHASE void ph_conv3T_pair(const LayerGeom g, int oZ) {
        const double* OUT = sm(oOUT); const double* W = sm(oW);
        double* C = sm(oZ);
        const int T = g.G * g.R, R = g.R, hR = R >> 1;
        for (int t2 = ex.tid(); t2 < (T >> 1); t2 += ex.nt()) {
            const int gi = t2 / hR, r = 2 * (t2 - gi * hR);
            const int rm = r == 0 ? R - 1 : r - 1, rp = r + 2 == R ? 0 : r + 2;
            double ob[NOUT][4];                              // out-gradient rows r-1 .. r+2; output row r + d reads row r + d - a + 1 = ob[.][d - a + 2]
#pragma unroll
            for (int o = 0; o < NOUT; ++o) {
                const double* p = OUT + o * T + gi * R;
                const dbl2 m = ld2(p + r);
                ob[o][0] = p[rm]; ob[o][1] = m.x; ob[o][2] = m.y; ob[o][3] = p[rp];
            }
#pragma unroll 1
            for (int k = 0; k < 3; ++k) {
                double acc[2][NH];
#pragma unroll
                for (int ci = 0; ci < NH; ++ci) { acc[0][ci] = 0.0; acc[1][ci] = 0.0; }
#pragma unroll
                for (int o = 0; o < NOUT; ++o)
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        double w[NH];
#pragma unroll
                        for (int ci = 0; ci < NH; ci += 2) { const dbl2 wv = ld2(W + OFF_W3T + ((o * 3 + a) * 3 + k) * NH + ci); w[ci] = wv.x; w[ci + 1] = wv.y; }
#pragma unroll
                        for (int ci = 0; ci < NH; ++ci) {
                            acc[0][ci] = fma(w[ci], ob[o][2 - a], acc[0][ci]);
                            acc[1][ci] = fma(w[ci], ob[o][3 - a], acc[1][ci]);
                        }
                    }
#pragma unroll
                for (int ci = 0; ci < NH; ++ci) {
                    double* p = C + ci * sB + (3 * gi + k) * R + r;
                    const dbl2 d = ld2(p);                   // act'(z2) of the two rows
                    st2(p, acc[0][ci] * d.x, acc[1][ci] * d.y);
                }
            }
        }
    }
    template <int CH> FT_PHASE void ph_conv3T_t(const LayerGeom g, int oZ) {
        const double* OUT = sm(oOUT); const double* W = sm(oW);
        double* C = sm(oZ);
        const int T = g.G * g.R, R = g.R;
        for (int t2 = ex.tid(); t2 < T * (NH / CH); t2 += ex.nt()) {
            int gi, h, r;
            task_split<NH / CH>(g, t2, gi, h, r);
            // out-gradient rows r-a+1 for a=0,1,2 -> r+1, r, r-1
            int rs[3] = { r + 1 == R ? 0 : r + 1, r, r == 0 ? R - 1 : r - 1 };
            double ob[NOUT][3];
#pragma unroll
            for (int o = 0; o < NOUT; ++o)
#pragma unroll
                for (int a = 0; a < 3; ++a) ob[o][a] = OUT[o * T + gi * R + rs[a]];
            double d2v[3][CH];                               // act'(z2) of this task's outputs, fetched before the MAC chains
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
                for (int ci = 0; ci < CH; ++ci) d2v[k][ci] = C[(CH * h + ci) * sB + (3 * gi + k) * R + r];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                double acc[CH];
#pragma unroll
                for (int ci = 0; ci < CH; ++ci) acc[ci] = 0.0;
#pragma unroll
                for (int o = 0; o < NOUT; ++o)
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        double w[CH];
#pragma unroll
                        for (int ci = 0; ci < CH; ci += 2) { const dbl2 wv = ld2(W + OFF_W3T + ((o * 3 + a) * 3 + k) * NH + CH * h + ci); w[ci] = wv.x; w[ci + 1] = wv.y; }
#pragma unroll
                        for (int ci = 0; ci < CH; ++ci) acc[ci] = fma(w[ci], ob[o][a], acc[ci]);
                    }
#pragma unroll
                for (int ci = 0; ci < CH; ++ci) {
                    int idx = (CH * h + ci) * sB + (3 * gi + k) * R + r;
                    C[idx] = acc[ci] * d2v[k][ci];
                }
            }
        }
    }

    // zbar1 = conv2^T(zbar2) * act'(z1)  (in place in A, which holds act'(z1)); same task shape as ph_conv2
    FT_HD void ph_conv2T(const LayerGeom g, int oZ) {
        if (wino_ok(g)) ph_conv2T_wino(g, oZ);
        else if (mma_ok()) ph_conv2T_mma(g, oZ);
        else if (fine_tasks()) ph_conv2T_t<2>(g, oZ);
        else ph_conv2T_t<4>(g, oZ);
    }

    // conv2^T on the fp64 tensor path: D[8 rows of output column 4g+q][8 input channels ci] += A[8 rows][4 channels o of
    // zbar2 at one tap] * B[o][ci].  One warp task = 8 rows of one stripe group, the four output columns (four accumulator
    // tiles); per (a, b) only the output columns whose source column 4g + q - b + 1 carries zbar2 take part (9 of 12).
    FT_PHASE void ph_conv2T_mma(const LayerGeom g, int oZ) {
        constexpr int NL = E::kLanes;
        const double* C = sm(oZ); const double* W = sm(oW);
        double* A = sm(oA);
        const int R = g.R, G = g.G, RB = R >> 3;
        double bf[18][NL];                                   // chunk c = tap * 2 + half: rows o = 4*half + j, columns ci = n
        FT_LANES(ln, ls) {
            const int j = ln & 3, n = ln >> 2;
#pragma unroll
            for (int c = 0; c < 18; ++c) bf[c][ls] = W[OFF_W2T + ((4 * (c & 1) + j) * 9 + (c >> 1)) * NH + n];
        }
        wait_bar(BAR_D1);                                    // act'(z1) has landed in A (issued a layer ago)
        for (int st = ex.warp(); st < G * RB; st += ex.nwarps()) {
            const int gi = st / RB, rb = 8 * (st - gi * RB);
            const int gn = gi + 1 == G ? 0 : gi + 1;
            // source column slots m = 0..4: (gi,k=0),(gi,1),(gi,2),(gn,0),(gn,1) == columns 4g-1, 4g, 4g+1, 4g+3, 4g+4
            const bool hal = CL && gi + 1 == G;              // cluster mode: slots 3, 4 of the last group sit in ZH[o][m-3][r]
            const int sc[5] = { 3 * gi * R, (3 * gi + 1) * R, (3 * gi + 2) * R,
                                hal ? (oC - oZ) : 3 * gn * R, hal ? (oC - oZ) + R : (3 * gn + 1) * R };
            double acc0[4][NL], acc1[4][NL], d1a[4][NL], d1b[4][NL];
            int rowa[3][NL];                                 // output row r reads source row r - a + 1
            int hcj[NL];
            FT_LANES(ln, ls) {
                const int i = ln >> 2, j = ln & 3, r = rb + i;
                hcj[ls] = CL ? j * (2 * R - sB) : 0;
                rowa[0][ls] = j * sB + (r + 1 == R ? 0 : r + 1);
                rowa[1][ls] = j * sB + r;
                rowa[2][ls] = j * sB + (r == 0 ? R - 1 : r - 1);
#pragma unroll
                for (int q = 0; q < 4; ++q) {                // act'(z1) of this lane's outputs: fetched under the DMMAs
                    const double* p = A + 2 * j * sA + (4 * gi + q) * R + rb + i;
                    acc0[q][ls] = 0.0; acc1[q][ls] = 0.0; d1a[q][ls] = p[0]; d1b[q][ls] = p[sA];
                }
            }
#ifdef FT_PROFILE
            long long tq0 = ex.clock();
#endif
#pragma unroll
            for (int c = 0; c < 18; ++c) {
                const int tap = c >> 1, a = tap / 3, b = tap - 3 * a, hoff = 4 * (c & 1) * sB;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int co = q - b + 1;                // source column relative to 4g: must be one of -1, 0, 1, 3, 4
                    const int m = co == -1 ? 0 : co == 0 ? 1 : co == 1 ? 2 : co == 3 ? 3 : co == 4 ? 4 : -1;
                    if (m >= 0) {
                        double av[NL];
                        const bool halo = hal && m >= 3;
                        const int hcorr = halo ? 4 * (c & 1) * (2 * R - sB) : 0;
                        FT_LANES(ln, ls) av[ls] = C[rowa[a][ls] + sc[m] + hoff + (halo ? hcj[ls] + hcorr : 0)];
                        ex.mma884(acc0[q], acc1[q], av, bf[c]);
                    }
                }
            }
#ifdef FT_PROFILE
            ex.prof_add(PF_C2T_MAC, ex.clock() - tq0);
#endif
            FT_LANES(ln, ls) {
                const int i = ln >> 2, j = ln & 3;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    double* p = A + 2 * j * sA + (4 * gi + q) * R + rb + i;
                    p[0] = acc0[q][ls] * d1a[q][ls];
                    p[sA] = acc1[q][ls] * d1b[q][ls];
                }
            }
        }
    }
    // conv2^T in the same Winograd F(2,3) form: output row r reads source row r - a + 1, i.e. a row kernel flipped
    // (g'_a' = W[2 - a']), so u' = (g2, (g0+g1+g2)/2, (g0-g1+g2)/2, g0).  One warp task = 8 row pairs of one stripe
    // group, the four output columns (sixteen accumulator tiles); the five source slots feed 1,2,3,2,1 of them.
    FT_PHASE void ph_conv2T_wino(const LayerGeom g, int oZ) {
        constexpr int NL = E::kLanes;
        const double* C = sm(oZ); const double* W = sm(oW);
        double* A = sm(oA);
        const int R = g.R, G = g.G, RB = R >> 4;
        double bf[3][2][4][NL];                              // rows o = 4*half + j, columns ci = n
        FT_LANES(ln, ls) {
            const int j = ln & 3, n = ln >> 2;
#pragma unroll
            for (int b = 0; b < 3; ++b)
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    const double* wp = W + OFF_W2T + ((4 * hf + j) * 9 + b) * NH + n;
                    const double g0 = wp[0], g1 = wp[3 * NH], g2 = wp[6 * NH], gs = g0 + g2;
                    bf[b][hf][0][ls] = g2; bf[b][hf][1][ls] = 0.5 * (gs + g1); bf[b][hf][2][ls] = 0.5 * (gs - g1); bf[b][hf][3][ls] = g0;
                }
        }
        wait_bar(BAR_D1);                                    // act'(z1) has landed in A (issued a layer ago)
        for (int st = ex.warp(); st < G * RB; st += ex.nwarps()) {
            const int gi = st / RB, rb = 16 * (st - gi * RB);
            const int gn = gi + 1 == G ? 0 : gi + 1;
            // source column slots s = 0..4: (gi,k=0),(gi,1),(gi,2),(gn,0),(gn,1) == columns 4g-1, 4g, 4g+1, 4g+3, 4g+4
            const bool hal = CL && gi + 1 == G;              // cluster mode: slots 3, 4 of the last group sit in ZH[o][s-3][r]
            const int sc[5] = { 3 * gi * R, (3 * gi + 1) * R, (3 * gi + 2) * R,
                                hal ? (oC - oZ) : 3 * gn * R, hal ? (oC - oZ) + R : (3 * gn + 1) * R };
            double acc0[4][4][NL], acc1[4][4][NL];
            int rm[NL], r1[NL], rp[NL];
            int hcj[NL];
            FT_LANES(ln, ls) {
                const int i = ln >> 2, j = ln & 3, r0 = rb + 2 * i;
                hcj[ls] = CL ? j * (2 * R - sB) : 0;
                rm[ls] = j * sB + (r0 == 0 ? R - 1 : r0 - 1);
                r1[ls] = j * sB + r0;
                rp[ls] = j * sB + (r0 + 2 == R ? 0 : r0 + 2);
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int m = 0; m < 4; ++m) { acc0[q][m][ls] = 0.0; acc1[q][m][ls] = 0.0; }
            }
#ifdef FT_PROFILE
            long long tq0 = ex.clock();
#endif
#pragma unroll
            for (int s5 = 0; s5 < 5; ++s5)
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    double t[4][NL];
                    const bool halo = hal && s5 >= 3;
                    FT_LANES(ln, ls) {
                        const double* pc = C + sc[s5] + 4 * hf * sB + (halo ? hcj[ls] + 4 * hf * (2 * R - sB) : 0);
                        const double d0 = pc[rm[ls]], d3 = pc[rp[ls]];
                        const dbl2 d12 = ld2(pc + r1[ls]);
                        t[0][ls] = d0 - d12.y; t[1][ls] = d12.x + d12.y; t[2][ls] = d12.y - d12.x; t[3][ls] = d12.x - d3;
                    }
                    const int co = s5 < 3 ? s5 - 1 : s5;     // source column relative to 4g: -1, 0, 1, 3, 4
#pragma unroll
                    for (int b = 0; b < 3; ++b) {
                        const int q = co + b - 1;            // the output column that sees this source through tap b
                        if (q >= 0 && q < 4) {
#pragma unroll
                            for (int m = 0; m < 4; ++m) ex.mma884(acc0[q][m], acc1[q][m], t[m], bf[b][hf][m]);
                        }
                    }
                }
#ifdef FT_PROFILE
            ex.prof_add(PF_C2T_MAC, ex.clock() - tq0);
#endif
            FT_LANES(ln, ls) {
                const int i = ln >> 2, j = ln & 3;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    double* p = A + 2 * j * sA + (4 * gi + q) * R + rb + 2 * i;
                    const dbl2 da = ld2(p), db = ld2(p + sA);     // act'(z1) of rows (2i, 2i+1), channels 2j and 2j+1
                    st2(p, ((acc0[q][0][ls] + acc0[q][1][ls]) + acc0[q][2][ls]) * da.x,
                           ((acc0[q][1][ls] - acc0[q][2][ls]) - acc0[q][3][ls]) * da.y);
                    st2(p + sA, ((acc1[q][0][ls] + acc1[q][1][ls]) + acc1[q][2][ls]) * db.x,
                                ((acc1[q][1][ls] - acc1[q][2][ls]) - acc1[q][3][ls]) * db.y);
                }
            }
        }
    }
    template <int CH> FT_PHASE void ph_conv2T_t(const LayerGeom g, int oZ) {
        const double* C = sm(oZ); const double* W = sm(oW);
        double* A = sm(oA);
        const int T = g.G * g.R * (4 / CH), R = g.R, G = g.G;
        for (int t = ex.tid(); t < T; t += ex.nt()) {
            int gi, h, r0;
            task2<CH>(g, t, gi, h, r0);
            const int gn = gi + 1 == G ? 0 : gi + 1;
            const int rm = r0 == 0 ? R - 1 : r0 - 1, rp = r0 + 2 == R ? 0 : r0 + 2;
            // source column slots j=0..4: (gi,k=0),(gi,1),(gi,2),(gn,0),(gn,1) == columns 4g-1,4g,4g+1,4g+3,4g+4
            // In cluster mode the two columns right of the rank's last group sit in the halo buffer ZH[o][j][r] (arena C;
            // the d2 prefetch is single-buffered in B there)
            const bool hal = CL && gi + 1 == G;
            const int sc[5] = { 3 * gi * R, (3 * gi + 1) * R, (3 * gi + 2) * R,
                                hal ? (oC - oZ) : 3 * gn * R, hal ? (oC - oZ) + R : (3 * gn + 1) * R };
            const int sst = hal ? 2 * R : sB;                  // channel stride of slots 3, 4
            const int CO[5] = { -1, 0, 1, 3, 4 };                                // column offsets from 4g
#ifdef FT_PROFILE
            long long tp0 = ex.clock();
#endif
            double acc[2][4][CH];                                                // [row][column q][channel]
#pragma unroll
            for (int dr = 0; dr < 2; ++dr)
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int ci = 0; ci < CH; ++ci) acc[dr][q][ci] = 0.0;
#pragma unroll 1
            for (int o = 0; o < NH; ++o) {
                const double* Cp = C + o * sB;
                double zb[4][5];                                                 // rows r0-1 .. r0+2
#pragma unroll
                for (int j = 0; j < 5; ++j) {
                    const double* col = (CL && j >= 3) ? C + sc[j] + o * sst : Cp + sc[j];
                    zb[0][j] = col[rm];
                    const dbl2 m = ld2(col + r0);
                    zb[1][j] = m.x; zb[2][j] = m.y;
                    zb[3][j] = col[rp];
                }
                const double* wo = W + OFF_W2T + o * 9 * NH + CH * h;
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int j = 0; j < 5; ++j)
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const int b = q - CO[j] + 1;           // source column = column - b + 1
                            if (b >= 0 && b <= 2) {
                                double w[CH];
#pragma unroll
                                for (int ci = 0; ci < CH; ci += 2) { const dbl2 wv = ld2(wo + (a * 3 + b) * NH + ci); w[ci] = wv.x; w[ci + 1] = wv.y; }
                                // output row r0+dr reads source row r0+dr-a+1 == zb[dr - a + 2]
#pragma unroll
                                for (int dr = 0; dr < 2; ++dr)
#pragma unroll
                                    for (int ci = 0; ci < CH; ++ci) acc[dr][q][ci] = fma(w[ci], zb[dr - a + 2][j], acc[dr][q][ci]);
                            }
                        }
            }
#ifdef FT_PROFILE
            ex.prof_add(PF_C2T_MAC, ex.clock() - tp0); tp0 = ex.clock();
#endif
            wait_bar(BAR_D1);                     // act'(z1) has landed in A
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int ci = 0; ci < CH; ++ci) {
                    double* p = A + (CH * h + ci) * sA + (4 * gi + q) * R + r0;
                    const dbl2 d = ld2(p);
                    st2(p, acc[0][q][ci] * d.x, acc[1][q][ci] * d.y);
                }
#ifdef FT_PROFILE
            ex.prof_add(PF_C3_CONV, ex.clock() - tp0);
#endif
        }
    }

    // (cos,sin)-gradients at the frozen sites = conv1^T(zbar1); assemble Pbar in the canonical layout
    // PB[c][r] (the unused forward-weight slots of W: V <= OFF_W3T doubles)
    FT_HD void ph_conv1T(const LayerGeom g) {
#if FT_CONV1T_PAIR
        if (!fine_tasks() && (g.R & 1) == 0) { ph_conv1T_pair(g); return; }
#endif
        if (fine_tasks()) ph_conv1T_t<1>(g); else ph_conv1T_t<2>(g);
    }
    // The same for the two rows (r, r+1), r even, of a stripe group per thread: every weight vector is loaded once for both
    // rows and the four input rows r-1 .. r+2 of a column are shared (as in conv3_pair) -- the phase is bound by shared-memory
    // wavefronts, and this form needs 41 of them per (channel, 2 rows) where the row-per-thread form needs 2 x 33.
    FT_PHASE void ph_conv1T_pair(const LayerGeom g) {
        const double* A = sm(oA); const double* W = sm(oW); const double* CS = sm(oCS); const double* UA = sm(oUA);
        double* PB = sm(oW);
        const int T = g.G * g.R, R = g.R, hR = R >> 1;
        for (int t2 = ex.tid(); t2 < (T >> 1); t2 += ex.nt()) {
            const int gi = t2 / hR, r = 2 * (t2 - gi * hR);
            const int rm = r == 0 ? R - 1 : r - 1, rp = r + 2 == R ? 0 : r + 2;
            // output row r + d reads source row r + d - a + 1: with v[j] = row r - 1 + j, that is v[d - a + 2]
            double gc[2][3][2], gs[2][3][2];                                      // [row d][kernel row a][frozen column k]
#pragma unroll
            for (int d = 0; d < 2; ++d)
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int k = 0; k < 2; ++k) { gc[d][a][k] = 0.0; gs[d][a][k] = 0.0; }
#pragma unroll 2
            for (int o = 0; o < NH; ++o) {
                double v[4][4];                                                   // [row r-1 .. r+2][column 4g .. 4g+3]
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double* col = A + o * sA + (4 * gi + q) * R;
                    const dbl2 m = ld2(col + r);
                    v[0][q] = col[rm]; v[1][q] = m.x; v[2][q] = m.y; v[3][q] = col[rp];
                }
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) {
                        const dbl2 w = ld2(W + OFF_W1T + ((o * 3 + a) * 3 + b) * 2);
#pragma unroll
                        for (int d = 0; d < 2; ++d)
#pragma unroll
                            for (int k = 0; k < 2; ++k) {                         // frozen column 4g+1+k reads column 4g+1+k-b+1
                                gc[d][a][k] = fma(w.x, v[d - a + 2][k + 2 - b], gc[d][a][k]);
                                gs[d][a][k] = fma(w.y, v[d - a + 2][k + 2 - b], gs[d][a][k]);
                            }
                    }
            }
            const dbl2 ua = ld2(UA + gi * R + r);
            st2(PB + (4 * gi) * R + r, ua.x, ua.y);
            st2(PB + (4 * gi + 3) * R + r, 0.0, 0.0);
            wait_bar(BAR_CS);                     // the frozen cos/sin have landed in CS
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const dbl2 cp = ld2(CS + (2 * gi + k) * R + r), sp = ld2(CS + V / 2 + (2 * gi + k) * R + r);
                const double c0 = (gc[0][0][k] + gc[0][1][k]) + gc[0][2][k], s0 = (gs[0][0][k] + gs[0][1][k]) + gs[0][2][k];
                const double c1 = (gc[1][0][k] + gc[1][1][k]) + gc[1][2][k], s1 = (gs[1][0][k] + gs[1][1][k]) + gs[1][2][k];
                st2(PB + (4 * gi + 1 + k) * R + r, -sp.x * c0 + cp.x * s0, -sp.y * c1 + cp.y * s1);
            }
        }
    }
    // KN = frozen columns per task: 2 (one task per (group, row)) or 1 (two tasks, wide blocks)
    template <int KN> FT_PHASE void ph_conv1T_t(const LayerGeom g) {
        const double* A = sm(oA); const double* W = sm(oW); const double* CS = sm(oCS); const double* UA = sm(oUA);
        double* PB = sm(oW);
        const int T = g.G * g.R, R = g.R;
        for (int t2 = ex.tid(); t2 < T * (2 / KN); t2 += ex.nt()) {
            int gi, k0, r;
            task_split<2 / KN>(g, t2, gi, k0, r);
            int rs[3] = { r + 1 == R ? 0 : r + 1, r, r == 0 ? R - 1 : r - 1 };   // r-a+1
            double gca[3][KN], gsa[3][KN];                                        // per kernel row: independent chains
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int k = 0; k < KN; ++k) { gca[a][k] = 0.0; gsa[a][k] = 0.0; }
#pragma unroll 2
            for (int o = 0; o < NH; ++o) {
                double v[3][KN + 2];                                              // rows r-a+1, columns 4g+k0 .. 4g+k0+KN+1
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int q = 0; q < KN + 2; ++q) v[a][q] = A[o * sA + (4 * gi + k0 + q) * R + rs[a]];
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) {
                        const dbl2 w = ld2(W + OFF_W1T + ((o * 3 + a) * 3 + b) * 2);
#pragma unroll
                        for (int k = 0; k < KN; ++k) {                            // frozen column 4g+1+k reads column 4g+1+k-b+1
                            gca[a][k] = fma(w.x, v[a][k + 2 - b], gca[a][k]);
                            gsa[a][k] = fma(w.y, v[a][k + 2 - b], gsa[a][k]);
                        }
                    }
            }
            if (k0 == 0) {
                PB[(4 * gi) * R + r] = UA[gi * R + r];
                PB[(4 * gi + 3) * R + r] = 0.0;
            }
            wait_bar(BAR_CS);                     // the frozen cos/sin have landed in CS
#pragma unroll
            for (int k = 0; k < KN; ++k) {
                const double gc = (gca[0][k] + gca[1][k]) + gca[2][k], gs = (gsa[0][k] + gsa[1][k]) + gsa[2][k];
                const double cp = CS[(2 * gi + k0 + k) * R + r], sp = CS[V / 2 + (2 * gi + k0 + k) * R + r];
                PB[(4 * gi + 1 + k0 + k) * R + r] = -sp * gc + cp * gs;
            }
        }
    }

    // GR += plaquette^T(Pbar); threads run along the stripe direction r (conflict-free on PB and on the padded GR)
    // Cluster mode: Pbar lives in canonical column blocks, GR in lattice row blocks.  Instead of read-modify-writing GR
    // through distributed shared memory (two dependent round trips per site), every rank sends its V values of Pbar to the
    // rank that owns the lattice row (remote stores into ST[H + 1][L1] in arena C, row 0 = the row above the block), and
    // after one cluster barrier each rank updates its own GR from its own shared memory: a quarter of the remote bytes, none
    // of them on a dependent path.  Same operands, same operations: bit-identical to the direct form.
    FT_PHASE void ph_scatter(const LayerGeom g) {
        if constexpr (CL) {
            const double* PB = sm(oW);
            const int R = g.R, oST = oStage();
            const int dcol = ex.nt() / R, drow = ex.nt() - dcol * R;
            int c = ex.tid() / R, r = ex.tid() - c * R;
            const int shH = log2_or_neg(H);
            for (int i = ex.tid(); i < V; i += ex.nt(), c += dcol, r += drow) {
                if (r >= R) { r -= R; ++c; }
                int n0, n1; site(g, r, c, n0, n1);
                const int j = shH >= 0 ? (n0 >> shH) : n0 / H, lr = n0 - j * H;
                const double pb = PB[i];
                ex.peer(sm(oST), j)[(lr + 1) * L1 + n1] = pb;
                if (lr == H - 1) ex.peer(sm(oST), j + 1 == nr ? 0 : j + 1)[n1] = pb;
            }
            ex.sync();
            const double* ST = sm(oST);
            double* GR = sm(oGR);
            const int e0 = ex.nt() / L1, e1 = ex.nt() - e0 * L1;
            int lr = ex.tid() / L1, n1 = ex.tid() - lr * L1;
            for (int i = ex.tid(); i < V; i += ex.nt(), lr += e0, n1 += e1) {
                if (n1 >= L1) { n1 -= L1; ++lr; }
                const int n1m = n1 == 0 ? L1 - 1 : n1 - 1;
                const double pb = ST[(lr + 1) * L1 + n1], pm1 = ST[(lr + 1) * L1 + n1m], pm0 = ST[lr * L1 + n1];
                GR[lr * LP + n1] += pb - pm1;
                GR[(H + lr) * LP + n1] += pm0 - pb;
            }
            return;
        }
        const double* PB = sm(oW);
        const int R = g.R, Cn = g.Cn;
        // (column, row) of the sites this thread visits advance without divisions: i += nt  <=>  c += nt / R, r += nt % R
        const int dcol = ex.nt() / R, drow = ex.nt() - dcol * R;
        int c = ex.tid() / R, r = ex.tid() - c * R;
        for (int i = ex.tid(); i < V; i += ex.nt(), c += dcol, r += drow) {
            if (r >= R) { r -= R; ++c; }
            // the column left of a rank's first column is the passive column of the previous group: Pbar == 0 there
            const int cm = c == 0 ? Cn - 1 : c - 1, rm = r == 0 ? R - 1 : r - 1;
            const double pb = PB[i], pmc = (CL && c == 0) ? 0.0 : PB[cm * R + r], pmr = PB[c * R + rm];
            int n0, n1; site(g, r, c, n0, n1);
            // mu=0: (n0,n1)=(r,c+off): P(n-e1)=pmc, P(n-e0)=pmr;  mu=1: (n0,n1)=(c+off,r): P(n-e1)=pmr, P(n-e0)=pmc
            const double pm1 = g.mu == 0 ? pmc : pmr, pm0 = g.mu == 0 ? pmr : pmc;
            *xat(oGR, 0, n0, n1) += pb - pm1;
            *xat(oGR, 1, n0, n1) += pm0 - pb;
        }
    }

    // ---- bulk (TMA) prefetch of the layer blocks for the adjoint sweep ----
    // Every block is contiguous in the workspace and lands at the same layout in shared memory, so each is ONE
    // cp.async.bulk issued by thread 0 and completed on its own transaction barrier: act'(z2) -> B (one layer ahead; with
    // -DFT_D2_SINGLE=0: B or C alternately, two layers ahead), act'(z1) -> A, frozen cos/sin -> CS, transposed weights -> W.  The waits sit
    // right before the first use (d1 / cs: after the MAC loops of ph_conv2T / ph_conv1T).
    // cluster mode: act'(z2) is single-buffered in B, one layer ahead; arena C holds the halos
    FT_HD int zbuf(int l) const { return (CL || FT_D2_SINGLE) ? oB : ((l & 1) ? oB : oC); }
    FT_HD int zbar(int l) const { return zbuf(l) == oB ? BAR_D2B : BAR_D2C; }
    FT_HD void issue_d2(int l) { if (l >= 0 && ex.tid() == 0) ex.bulk_load(zbar(l), sm(zbuf(l)), wsD2(l), NH * sB); }
    FT_HD void issue_d1(int l) { if (l >= 0 && ex.tid() == 0) ex.bulk_load(BAR_D1, sm(oA), wsD1(l), NH * sA); }
    FT_HD void issue_cs(int l) { if (l >= 0 && ex.tid() == 0) ex.bulk_load(BAR_CS, sm(oCS), wsCS(l), V); }
    // (s_1, s_2) and the pre-update active links: contiguous in the layer block, same order as OUT
    FT_HD void issue_so(int l) { if (l >= 0 && ex.tid() == 0) ex.bulk_load(BAR_SO, sm(oOUT), wsSO(l), 3 * VQ); }

    // Training-mode adjoint of one layer: the input-gradient phases plus the weight-gradient GEMMs, with the saved
    // activations staged through shared memory.  Plane use: A = h1(l) for ph_wgrad2, then act'(z1)(l) for ph_conv2T;
    // B = act'(z2)(l) (single-buffered); C = h2(l) for ph_wgrad3.  Every copy has its own transaction barrier and is issued
    // as soon as its destination is free; only act'(z1), which has to wait for ph_wgrad2 to release A, is exposed.
    FT_HD void issue_plane(int bar, int off, const double* src, int n) { if (ex.tid() == 0) ex.bulk_load(bar, sm(off), src, n); }
    FT_HD void layer_adjoint_train(int l) {
        LayerGeom g = geom(l);
        ph_outgrad(g);                                        // waits for so/sv(l)
        wait_bar(BAR_D2B); wait_bar(BAR_W);
        ex.lsync();
        advance_bar(BAR_D2B); advance_bar(BAR_W); advance_bar(BAR_SO);
        wait_bar(BAR_D2C);                                    // h2(l) has landed in C
        ph_wgrad3(g, sm(oC), gslice(l));
        ph_conv3T(g, oB);
        ex.lsync();
        advance_bar(BAR_D2C);
        issue_so(l - 1);                                      // OUT and C are free
        if (l >= 1) issue_plane(BAR_D2C, oC, wsH2(l - 1), NH * sB);
        wait_bar(BAR_H1);                                     // h1(l) has landed in A
        ph_wgrad2(g, oB, sm(oA), gslice(l));
        ex.lsync();
        advance_bar(BAR_H1);
        issue_plane(BAR_D1, oA, wsD1(l), NH * sA);            // act'(z1)(l) replaces h1(l)
        ph_conv2T(g, oB);
        ex.lsync();
        advance_bar(BAR_D1);
        if (l >= 1) issue_plane(BAR_D2B, oB, wsD2(l - 1), NH * sB);
        ph_wgrad1(g, gslice(l));
        ph_conv1T(g);
        ex.lsync();
        advance_bar(BAR_CS);
        issue_weights(l - 1, true);
        if (l >= 1) issue_plane(BAR_H1, oA, wsH1(l - 1), NH * sA);
        issue_cs(l - 1);
        ph_scatter(g);
        ex.sync();
    }

    // in flight on entry: d2(l) [, d2(l-1)], Wt(l), d1(l), cs(l)
    template <bool TRAIN = false> FT_HD void layer_adjoint(int l) {
        if (TRAIN) { layer_adjoint_train(l); return; }
        LayerGeom g = geom(l);
        FT_T(PF_OUTGRAD, if (CL && FT_CL_SAVED_U && nr > 1) ph_outgrad_saved(g, wsSV(l) + VQ); else ph_outgrad(g);   // waits for so/sv(l)
             wait_bar(zbar(l)); wait_bar(BAR_W);   // d2(l), Wt(l) have landed
             ex.lsync();                           // (restored links / GR reads cross ranks, but are ordered by the cluster barriers around)
             advance_bar(zbar(l)); advance_bar(BAR_W); advance_bar(BAR_SO));
        FT_T(PF_CONV3T, ph_conv3T(g, zbuf(l)); ex.lsync());
        FT_T(PF_ISSUE, issue_so(l - 1));           // OUT is free again
        if constexpr (CL) { push_halo_zbar2(g, zbuf(l)); ex.sync(); }
        FT_T(PF_CONV2T, ph_conv2T(g, zbuf(l));     // waits for d1(l) after its MAC loop
             ex.lsync();
             advance_bar(BAR_D1));
        FT_T(PF_ISSUE, issue_d2((CL || FT_D2_SINGLE) ? l - 1 : l - 2));   // zbuf(l) is free again
        FT_T(PF_CONV1T, ph_conv1T(g);              // waits for cs(l) after its MAC loop
             ex.lsync();
             advance_bar(BAR_CS));
        FT_T(PF_ISSUE, issue_weights(l - 1, true); // W(transposed), A and CS are free
             issue_d1(l - 1);
             issue_cs(l - 1));
        FT_T(PF_SCATTER, ph_scatter(g); ex.sync());
    }

    // =============================================================================================
    // chain programs on the resident field X
    // =============================================================================================
    // X <- F(X); returns sum of logJ (if asked).  save: write the layer blocks for the reverse sweep
    template <bool TRAIN = false> FT_HD double flow_forward(bool want_logJ, bool save, double* layer_logJ = nullptr) {
        double tot = 0.0;
        for (int l = 0; l < pr.nlayers; ++l) {
            double lj = layer_forward<TRAIN>(l, want_logJ, save);
            tot += lj;
            if (layer_logJ && ex.tid() == 0) layer_logJ[l] = lj;
        }
        return tot;
    }
    FT_HD double flow_reverse(bool want_logJ, double* layer_logJ = nullptr) {
        double tot = 0.0;
        for (int l = pr.nlayers - 1; l >= 0; --l) {
            double lj = layer_reverse(l, want_logJ);
            tot += lj;
            if (layer_logJ && ex.tid() == 0) layer_logJ[l] = lj;
        }
        return tot;
    }
    // ft_action (ipynb/ft_hmc.py:230-238): X <- F(X), returns S(F(x)) - sum logJ; *s_plain = S(F(x))
    FT_HD double ft_action(double beta, double* s_plain = nullptr) {
        double lj = flow_forward(true, false);
        double s = wilson_action(beta, pr.conv);
        if (s_plain) *s_plain = s;
        return s - lj;
    }
    // ft_force (ipynb/ft_hmc.py:240-249): GR <- d/dx [S(F(x)) - sum logJ]; X is preserved.
    // want_logJ (training): the forward sweep also accumulates sum logJ and ft_action(x) = S(F(x)) - sum logJ is returned
    template <bool TRAIN = false> FT_HD double ft_force(double beta, bool want_logJ = false) {
        double lj = flow_forward<TRAIN>(want_logJ, true);
        if (want_logJ) lj = wilson_action(beta, pr.conv) - lj;          // ft_action(x) = S(F(x)) - sum logJ
        ex.proxy_fence();                 // the layer blocks just written are read back by bulk copies (async proxy)
        ex.sync();
        const int last = pr.nlayers - 1;
        if (TRAIN) {
            issue_plane(BAR_D2B, oB, wsD2(last), NH * sB);
            issue_weights(last, true);
            issue_plane(BAR_H1, oA, wsH1(last), NH * sA);
            issue_plane(BAR_D2C, oC, wsH2(last), NH * sB);
            issue_cs(last);
        } else {
            FT_T(PF_ISSUE, issue_d2(last);
                 issue_d2((CL || FT_D2_SINGLE) ? -1 : last - 1);
                 issue_weights(last, true);
                 issue_d1(last);
                 issue_cs(last));
        }
        if (TRAIN && vjp_seed != nullptr) {                // vector-Jacobian mode: the caller's d/dy seeds the sweep
            double* GRs = sm(oGR);
            const double* sd = vjp_seed;
            for_links([&](int si, int gi) { GRs[si] = sd[gi]; });
            ex.sync();
        } else
        FT_T(PF_WFORCE, wilson_force(beta, pr.conv));      // scratch plane = UA+OUT
        FT_T(PF_ISSUE, issue_so(last));
        for (int l = last; l >= 0; --l) layer_adjoint<TRAIN>(l);
        return lj;
    }

    // =============================================================================================
    // weight gradients (flow training, ipynb/ft_hmc.py:253-295: d/dweights of sum_b [S(F(x_b)) - sum logJ]), accumulated
    // next to the input-gradient sweep.  Each is a GEMM with K = sites on the fp64 tensor path:
    //     D[8 output channels o][8 columns] += A[o][4 consecutive rows of one plane column] * B[those 4 sites][columns]
    // with the adjoint signal (OUT, zbar2, zbar1; shared memory) as A and the saved activations (h2, h1: bulk-copied from
    // the layer block into the planes C and A by layer_adjoint_train; cos/sin planes) as B.  One accumulator tile per kernel tap.
    // Every warp sums its share of the K chunks in registers and adds the result to ITS OWN slice of the CTA's gradient
    // buffer (no atomics; the host-side reduction over (CTA, warp) slices is in a fixed order).
    // =============================================================================================
    FT_HD double* gslice(int l) const { return gW + ((size_t)ex.warp() * pr.nlayers + l) * GRAD_DOUBLES; }
    FT_HD int wrapr(int r, int R) const { return r < 0 ? r + R : (r >= R ? r - R : r); }

    // conv3: dW3[o][ci][a][b] = sum_t OUT[o][t] h2[ci][3g+b][r+a-1], db3[o] = sum_t OUT[o][t]   (t = active sites)
    FT_PHASE void ph_wgrad3(const LayerGeom g, const double* h2g, double* gl) {
        constexpr int NL = E::kLanes;
        const double* OUT = sm(oOUT);
        const int T = g.G * g.R, R = g.R, RQ = R >> 2;
        double acc0[10][NL], acc1[10][NL];
#pragma unroll
        for (int i = 0; i < 10; ++i) FT_LANES(ln, ls) { acc0[i][ls] = 0.0; acc1[i][ls] = 0.0; }
        for (int ch = ex.warp(); ch < g.G * RQ; ch += ex.nwarps()) {
            const int gi = ch / RQ, r0 = 4 * (ch - gi * RQ);
            double av[NL], one[NL];
            FT_LANES(ln, ls) {
                const int o = ln >> 2, sI = ln & 3;
                av[ls] = o < NOUT ? OUT[o * T + gi * R + r0 + sI] : 0.0;
                one[ls] = (ln >> 2) == 0 ? 1.0 : 0.0;
            }
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const int a = tap / 3, b = tap - 3 * a;
                double bv[NL];
                FT_LANES(ln, ls) { const int sI = ln & 3, ci = ln >> 2; bv[ls] = h2g[ci * sB + (3 * gi + b) * R + wrapr(r0 + sI + a - 1, R)]; }
                ex.mma884(acc0[tap], acc1[tap], av, bv);
            }
            ex.mma884(acc0[9], acc1[9], av, one);
        }
        FT_LANES(ln, ls) {
            const int o = ln >> 2, j = ln & 3;
            if (o < NOUT) {
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    gl[OFF_W3F + ((2 * j) * 9 + tap) * 4 + o] += acc0[tap][ls];
                    gl[OFF_W3F + ((2 * j + 1) * 9 + tap) * 4 + o] += acc1[tap][ls];
                }
                if (j == 0) gl[OFF_B3 + o] += acc0[9][ls];
            }
        }
    }

    // conv2: dW2[o][ci][a][b] = sum z2bar[o][3g+k][r] h1[ci][4g-1+k+b-1][r+a-1], db2[o] = sum z2bar[o]
    FT_PHASE void ph_wgrad2(const LayerGeom g, int oZ, const double* h1g, double* gl) {
        constexpr int NL = E::kLanes;
        const double* Z = sm(oZ);
        const int R = g.R, Cn = g.Cn, RQ = R >> 2;
        double acc0[10][NL], acc1[10][NL];
#pragma unroll
        for (int i = 0; i < 10; ++i) FT_LANES(ln, ls) { acc0[i][ls] = 0.0; acc1[i][ls] = 0.0; }
        for (int ch = ex.warp(); ch < g.G * 3 * RQ; ch += ex.nwarps()) {
            const int gi = ch / (3 * RQ), rem = ch - gi * 3 * RQ, k = rem / RQ, r0 = 4 * (rem - k * RQ);
            double av[NL], one[NL];
            FT_LANES(ln, ls) {
                const int o = ln >> 2, sI = ln & 3;
                av[ls] = Z[o * sB + (3 * gi + k) * R + r0 + sI];
                one[ls] = (ln >> 2) == 0 ? 1.0 : 0.0;
            }
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const int a = tap / 3, b = tap - 3 * a;
                int c = 4 * gi + k + b - 2; c = c < 0 ? c + Cn : (c >= Cn ? c - Cn : c);
                double bv[NL];
                FT_LANES(ln, ls) { const int sI = ln & 3, ci = ln >> 2; bv[ls] = h1g[ci * sA + c * R + wrapr(r0 + sI + a - 1, R)]; }
                ex.mma884(acc0[tap], acc1[tap], av, bv);
            }
            ex.mma884(acc0[9], acc1[9], av, one);
        }
        FT_LANES(ln, ls) {
            const int o = ln >> 2, j = ln & 3;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                gl[OFF_W2F + ((2 * j) * 9 + tap) * NH + o] += acc0[tap][ls];
                gl[OFF_W2F + ((2 * j + 1) * 9 + tap) * NH + o] += acc1[tap][ls];
            }
            if (j == 0) gl[OFF_B2 + o] += acc0[9][ls];
        }
    }

    // conv1: output column class q sees the frozen column k through kernel column b = k + 2 - q:
    //   dW1[o][ci][a][b] += sum_{g,r} z1bar[o][4g+q][r] (cos,sin)[ci][2g+k][r+a-1]      (one tile per b, columns n = 2a + ci)
    // and the column-class sums S_q[o] = sum z1bar[o][class q] (tile columns 6, 7: ones for k = 0, 1), from which the
    // host derives db1 and the gradient through the constant (cos 0, sin 0) = (1, 0) inputs of the non-frozen sites.
    FT_PHASE void ph_wgrad1(const LayerGeom g, double* gl) {
        constexpr int NL = E::kLanes;
        const double* A = sm(oA); const double* CS = sm(oCS);
        const int R = g.R, RQ = R >> 2;
        wait_bar(BAR_CS);                                    // the frozen cos/sin of this layer have landed
        double acc0[3][NL], acc1[3][NL];
#pragma unroll
        for (int i = 0; i < 3; ++i) FT_LANES(ln, ls) { acc0[i][ls] = 0.0; acc1[i][ls] = 0.0; }
        for (int ch = ex.warp(); ch < 2 * g.G * RQ; ch += ex.nwarps()) {
            const int k = ch / (g.G * RQ), rem = ch - k * g.G * RQ, gi = rem / RQ, r0 = 4 * (rem - gi * RQ);
            double bv[NL];
            FT_LANES(ln, ls) {
                const int sI = ln & 3, n = ln >> 2;
                if (n < 6) bv[ls] = CS[(n & 1) * (V / 2) + (2 * gi + k) * R + wrapr(r0 + sI + (n >> 1) - 1, R)];
                else bv[ls] = (n - 6) == k ? 1.0 : 0.0;
            }
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                const int q = k + 2 - b;
                double av[NL];
                FT_LANES(ln, ls) { const int o = ln >> 2, sI = ln & 3; av[ls] = A[o * sA + (4 * gi + q) * R + r0 + sI]; }
                ex.mma884(acc0[b], acc1[b], av, bv);
            }
        }
        FT_LANES(ln, ls) {
            const int o = ln >> 2, j = ln & 3;
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                if (j < 3) {                                 // columns n = 2j, 2j+1  ->  kernel row a = j, ci = 0, 1
                    gl[OFF_W1F + ((b * 3 + j) * 2 + 0) * NH + o] += acc0[b][ls];
                    gl[OFF_W1F + ((b * 3 + j) * 2 + 1) * NH + o] += acc1[b][ls];
                } else if (b != 1) {                         // columns 6, 7: class sums; b = 0 -> S_2, S_3;  b = 2 -> S_0, S_1
                    gl[OFF_B1 + (2 - b) * NH + o] += acc0[b][ls];
                    gl[OFF_B1 + (3 - b) * NH + o] += acc1[b][ls];
                }
            }
        }
    }

    // elementwise helpers on the link field (skip the pitch padding)
    template <class F> FT_HD void for_links(F f) {
        for (int i = ex.tid(); i < 2 * V; i += ex.nt()) {
            int si, gi; link_map(i, si, gi);
            f(si, gi);
        }
    }
};

// ------------------------------------------------------------------------------------------------
// trajectory programs (one chain).  en.wsP / wsX0 / wsY0 (momenta, x0, y0) point into shared memory (plane C) for single-CTA
// flow chains and into the per-CTA global workspace otherwise (see the Engine constructor).
// ------------------------------------------------------------------------------------------------
struct TrajIO {
    const double* field_in;    // (2,L0,L1)
    double* field_out;         // (2,L0,L1)
    const double* p_in;        // (2,L0,L1) or null -> Philox
    const double* u_in;        // scalar or null -> Philox
    double* p_out;             // optional final momenta
    uint64_t seed, chain, traj;
    double beta, dt; int nstep;
    double* out_dH; double* out_expmdH; int* out_acc; double* out_plaq; double* out_Q;
    double* out_h0; double* out_h1;
    // multi-trajectory runs keep the chain resident: only the first trajectory of a launch loads the field from
    // global memory and only the last one stores it
    bool first, last;
};

// Box-Muller normals from Philox: element pair index j -> two normals
template <class E>
FT_HD void philox_momenta(Engine<E>& en, const TrajIO& io, double* P) {
    Philox ph{ (uint32_t)io.seed, (uint32_t)(io.seed >> 32) };
    // each rank draws the momenta of its own links; the counter is the global pair index, so the stream does
    // not depend on the decomposition (pairs never straddle a lattice row: L1 is even)
    for (int jl = en.ex.tid(); jl < en.V; jl += en.ex.nt()) {
        int si, gi; en.link_map(2 * jl, si, gi);
        const int j = gi >> 1;
        uint32_t r[4];
        ph.gen((uint32_t)j, (uint32_t)io.traj, (uint32_t)io.chain, (uint32_t)(io.chain >> 32) ^ 0x5EEDu, r);
        double u1 = u53(r[0], r[1]), u2 = u53(r[2], r[3]);
        double rad = sqrt(-2.0 * log(u1)), sn, cs;
        sn = sin(TWO_PI_D * u2); cs = cos(TWO_PI_D * u2);
        P[2 * j] = rad * cs; P[2 * j + 1] = rad * sn;
    }
}
FT_HD double philox_uniform(const TrajIO& io) {
    Philox ph{ (uint32_t)io.seed, (uint32_t)(io.seed >> 32) };
    uint32_t r[4];
    ph.gen(0xFFFFFFFFu, (uint32_t)io.traj, (uint32_t)io.chain, (uint32_t)(io.chain >> 32) ^ 0xACCE97u, r);
    return u53(r[0], r[1]);
}

// leapfrog skeleton shared by the plain and the field-transformed trajectory
// (hmc_2dU1.py:132-141, ipynb/ft_hmc.py:394-418): position-first, nstep force evaluations.
template <class E, class ForceFn>
FT_HD void leapfrog_resident(Engine<E>& en, double dt, int nstep, double* P, ForceFn force) {
    auto& ex = en.ex;
    const double hdt = 0.5 * dt;
    double* X = en.sm(en.oX);
    const double* GR = en.sm(en.oGR);
    en.for_links([&](int si, int gi) { X[si] = X[si] + hdt * P[gi]; });
    ex.sync();
    for (int s = 0; s < nstep; ++s) {
        force();
        const bool last = s == nstep - 1;
        const double step = last ? hdt : dt;
        en.for_links([&](int si, int gi) {
            double pn = P[gi] + (-dt) * GR[si];
            P[gi] = pn;
            X[si] = X[si] + step * pn;
        });
        ex.sync();
    }
}

// the MD steps of leapfrog_plain_fused for threads that own at most NS sites each (k_chain_plain): see there
template <class E, int NS>
FT_HD void plain_md_steps(Engine<E>& en, double beta, double dt, int nstep) {
    auto& ex = en.ex;
    const int L0 = en.L0, L1 = en.L1, LP = en.LP, V = en.V, X1 = L0 * LP;
    const double hdt = 0.5 * dt;
    double* X = en.sm(en.oX);
    double* P = en.sm(en.oGR);
    double* S = en.sm(en.oS);
    int xb[NS], sb[NS], dc[NS], dr[NS], sl[NS], su[NS];   // x0 offset, S offset, offsets of the n1+1 / n0+1 / n1-1 / n0-1 neighbours
    bool ok[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        const int i = ex.tid() + k * ex.nt();
        ok[k] = i < V;
        const int n0 = ok[k] ? i / L1 : 0, n1 = ok[k] ? i - n0 * L1 : 0;
        xb[k] = n0 * LP + n1; sb[k] = n0 * L1 + n1;
        dc[k] = n1 + 1 == L1 ? 1 - L1 : 1;
        dr[k] = n0 + 1 == L0 ? -(L0 - 1) * LP : LP;
        sl[k] = n1 == 0 ? L1 - 1 : -1;
        su[k] = n0 == 0 ? V - L1 : -L1;
    }
    for (int st = 0; st < nstep; ++st) {
        const double step = st == nstep - 1 ? hdt : dt;
        // the plaquettes of the thread's sites, ONE range test for all of them, then NS independent sine chains (a branch per
        // site put NS convergence regions in the step and serialised the chains)
        double pl[NS];
        bool inr = FT_FAST_SIN != 0;
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            pl[k] = 0.0;
            if (ok[k]) {
                const double a = X[xb[k]], b = X[X1 + xb[k] + dr[k]], c = X[xb[k] + dc[k]], d = X[X1 + xb[k]];
                pl[k] = ((a - d) - c) + b;
            }
            inr = inr && fabs(pl[k]) < 524288.0;
        }
        if (inr) {
#pragma unroll
            for (int k = 0; k < NS; ++k) pl[k] = sin_core(pl[k]);
        } else {
#pragma unroll
            for (int k = 0; k < NS; ++k) pl[k] = sin_force(pl[k]);
        }
#pragma unroll
        for (int k = 0; k < NS; ++k)
            if (ok[k]) S[sb[k]] = pl[k];
        ex.sync();
#pragma unroll
        for (int k = 0; k < NS; ++k)
            if (ok[k]) {
                const double sv = S[sb[k]];
                const double f0 = beta * (sv - S[sb[k] + sl[k]]), f1 = beta * (S[sb[k] + su[k]] - sv);
                const int i0 = xb[k], i1 = X1 + xb[k];
                const double p0 = P[i0] + (-dt) * f0, p1 = P[i1] + (-dt) * f1;
                P[i0] = p0; P[i1] = p1;
                X[i0] = X[i0] + step * p0;
                X[i1] = X[i1] + step * p1;
            }
        ex.sync();
    }
}

// Plain-HMC leapfrog (hmc_2dU1.py:132-141) with everything on chip: the momenta live in the (otherwise unused) gradient
// plane, and every MD step is two phases -- S = sin(plaquette), then per SITE the force of its two links from S, the
// momentum update and the position update -- with division-free site stepping.  Same operations in the same order as
// leapfrog_resident + wilson_force (bit-identical results), without the per-step trips of the momenta through L2, the
// separate gradient plane and a third barrier.  Single-CTA chains only; clusters use the generic path.
// FAST: the dedicated sine (k_chain_plain); the flow kernels, which carry this program only for completeness, keep the library's
template <class E, bool FAST = false>
FT_HD void leapfrog_plain_fused(Engine<E>& en, double beta, double dt, int nstep, double* Pg) {
    auto& ex = en.ex;
    const int L0 = en.L0, L1 = en.L1, LP = en.LP, V = en.V;
    const double hdt = 0.5 * dt;
    double* X = en.sm(en.oX);
    double* P = en.sm(en.oGR);                               // momenta, same pitch as X
    double* S = en.sm(en.oS);
    const int d0 = ex.nt() / L1, d1 = ex.nt() - d0 * L1, s0 = ex.tid() / L1, s1 = ex.tid() - s0 * L1;
    // rows of the (2 L0, L1) link matrix: row = mu * L0 + n0
    for (int i = ex.tid(), row = s0, n1 = s1; i < 2 * V; i += ex.nt(), row += d0, n1 += d1) {
        if (n1 >= L1) { n1 -= L1; ++row; }
        const double p = Pg[i];
        P[row * LP + n1] = p;
        X[row * LP + n1] = X[row * LP + n1] + hdt * p;
    }
    ex.sync();
    if constexpr (FAST) {
        // k_chain_plain, up to four sites per thread (L <= 32 with 256 threads): the kernel is ISSUE bound (ncu: issue slots
        // 70 % busy, fp64 pipe 25 %) and most of what it issued was the address arithmetic of the sites -- which are the same
        // in every MD step of every chain.  Offsets and wrap flags of a thread's sites are formed once; a step is then
        // loads, the plaquette, the sine, and the updates.  Same operands and operations as the loops below: bit-identical.
        if (V <= 4 * ex.nt()) {
            if (V <= ex.nt()) plain_md_steps<E, 1>(en, beta, dt, nstep);
            else if (V <= 2 * ex.nt()) plain_md_steps<E, 2>(en, beta, dt, nstep);
            else plain_md_steps<E, 4>(en, beta, dt, nstep);
            for (int i = ex.tid(), row = s0, n1 = s1; i < 2 * V; i += ex.nt(), row += d0, n1 += d1) {
                if (n1 >= L1) { n1 -= L1; ++row; }
                Pg[i] = P[row * LP + n1];
            }
            ex.sync();
            return;
        }
    }
    for (int st = 0; st < nstep; ++st) {
        const double step = st == nstep - 1 ? hdt : dt;
        for (int i = ex.tid(), n0 = s0, n1 = s1; i < V; i += ex.nt(), n0 += d0, n1 += d1) {
            if (n1 >= L1) { n1 -= L1; ++n0; }
            S[i] = FAST ? sin_force(en.plaq(en.oX, n0, n1, 1)) : sin(en.plaq(en.oX, n0, n1, 1));
        }
        ex.sync();
        for (int i = ex.tid(), n0 = s0, n1 = s1; i < V; i += ex.nt(), n0 += d0, n1 += d1) {
            if (n1 >= L1) { n1 -= L1; ++n0; }
            const int n0m = n0 == 0 ? L0 - 1 : n0 - 1, n1m = n1 == 0 ? L1 - 1 : n1 - 1;
            const double sv = S[i];
            const double f0 = beta * (sv - S[n0 * L1 + n1m]), f1 = beta * (S[n0m * L1 + n1] - sv);
            const int i0 = n0 * LP + n1, i1 = (L0 + n0) * LP + n1;
            const double p0 = P[i0] + (-dt) * f0, p1 = P[i1] + (-dt) * f1;
            P[i0] = p0; P[i1] = p1;
            X[i0] = X[i0] + step * p0;
            X[i1] = X[i1] + step * p1;
        }
        ex.sync();
    }
    for (int i = ex.tid(), row = s0, n1 = s1; i < 2 * V; i += ex.nt(), row += d0, n1 += d1) {
        if (n1 >= L1) { n1 -= L1; ++row; }
        Pg[i] = P[row * LP + n1];
    }
    ex.sync();
}

// The same leapfrog with the momenta already in, and staying in, the gradient plane (hmc_trajectory_plain): no copies
template <class E>
FT_HD void leapfrog_plain_resident(Engine<E>& en, double beta, double dt, int nstep) {
    auto& ex = en.ex;
    const int L0 = en.L0, L1 = en.L1, LP = en.LP, V = en.V;
    const double hdt = 0.5 * dt;
    double* X = en.sm(en.oX);
    double* P = en.sm(en.oGR);
    double* S = en.sm(en.oS);
    en.for_links([&](int si, int) { X[si] = X[si] + hdt * P[si]; });
    ex.sync();
    if (V <= ex.nt()) { plain_md_steps<E, 1>(en, beta, dt, nstep); return; }
    if (V <= 2 * ex.nt()) { plain_md_steps<E, 2>(en, beta, dt, nstep); return; }
    if (V <= 4 * ex.nt()) { plain_md_steps<E, 4>(en, beta, dt, nstep); return; }
    const int d0 = ex.nt() / L1, d1 = ex.nt() - d0 * L1, s0 = ex.tid() / L1, s1 = ex.tid() - s0 * L1;
    for (int st = 0; st < nstep; ++st) {
        const double step = st == nstep - 1 ? hdt : dt;
        for (int i = ex.tid(), n0 = s0, n1 = s1; i < V; i += ex.nt(), n0 += d0, n1 += d1) {
            if (n1 >= L1) { n1 -= L1; ++n0; }
            S[i] = sin_force(en.plaq(en.oX, n0, n1, 1));
        }
        ex.sync();
        for (int i = ex.tid(), n0 = s0, n1 = s1; i < V; i += ex.nt(), n0 += d0, n1 += d1) {
            if (n1 >= L1) { n1 -= L1; ++n0; }
            const int n0m = n0 == 0 ? L0 - 1 : n0 - 1, n1m = n1 == 0 ? L1 - 1 : n1 - 1;
            const double sv = S[i];
            const double f0 = beta * (sv - S[n0 * L1 + n1m]), f1 = beta * (S[n0m * L1 + n1] - sv);
            const int i0 = n0 * LP + n1, i1 = (L0 + n0) * LP + n1;
            const double p0 = P[i0] + (-dt) * f0, p1 = P[i1] + (-dt) * f1;
            P[i0] = p0; P[i1] = p1;
            X[i0] = X[i0] + step * p0;
            X[i1] = X[i1] + step * p1;
        }
        ex.sync();
    }
}

// FT-HMC trajectory (ipynb/ft_hmc.py:420-435)
template <class E>
FT_HD void ft_hmc_trajectory(Engine<E>& en, const TrajIO& io) {
    auto& ex = en.ex;
    const int V = en.Vg;
    double* P = en.wsP;
    double* X = en.sm(en.oX);
    if (io.first) en.load_field(en.oX, io.field_in);
    ex.sync();
    en.flow_reverse(false);                                     // x = ft_flow_inv(field)
    en.for_links([&](int si, int gi) { en.wsX0[gi] = X[si]; });
    if (io.p_in) en.for_links([&](int, int gi) { P[gi] = io.p_in[gi]; });
    else philox_momenta(en, io, P);
    ex.sync();
    double k0 = 0.0;
    en.for_links([&](int, int gi) { k0 += P[gi] * P[gi]; });
    k0 = ex.sum(k0);
    double s0_plain;
    double h0 = en.ft_action(io.beta, &s0_plain) + 0.5 * k0;    // X <- y0 = F(x)
    en.for_links([&](int si, int gi) { en.wsY0[gi] = X[si]; });
    ex.sync();
    en.for_links([&](int si, int gi) { X[si] = en.wsX0[gi]; });
    ex.sync();
    leapfrog_resident(en, io.dt, io.nstep, P, [&]() { en.ft_force(io.beta); });
    en.for_links([&](int si, int gi) { X[si] = regularize1(X[si]); });
    ex.sync();
    double k1 = 0.0;
    en.for_links([&](int, int gi) { k1 += P[gi] * P[gi]; });
    k1 = ex.sum(k1);
    double s1_plain;
    double h1 = en.ft_action(io.beta, &s1_plain) + 0.5 * k1;    // X <- F(xr)
    double u = io.u_in ? *io.u_in : philox_uniform(io);
    double dH = h1 - h0;
    double e = exp(-dH);
    bool acc = u < e;
    if (!acc) {                                                  // newfield = ft_flow(x) = y0
        ex.sync();
        en.for_links([&](int si, int gi) { X[si] = en.wsY0[gi]; });
    }
    ex.sync();
    double q = en.topo_floor();
    if (io.last) en.store_field(io.field_out, en.oX);
    if (io.p_out) en.for_links([&](int, int gi) { io.p_out[gi] = P[gi]; });
    if (ex.tid() == 0 && en.rk == 0) {
        if (io.out_dH) *io.out_dH = dH;
        if (io.out_expmdH) *io.out_expmdH = e;
        if (io.out_acc) *io.out_acc = acc ? 1 : 0;
        if (io.out_plaq) *io.out_plaq = (acc ? s1_plain : s0_plain) / (-io.beta * V);
        if (io.out_Q) *io.out_Q = q;
        if (io.out_h0) *io.out_h0 = h0;
        if (io.out_h1) *io.out_h1 = h1;
    }
    ex.sync();
}

// Plain HMC trajectory of k_chain_plain (single-CTA chains): the program of hmc_trajectory below with everything that is not
// the reference's arithmetic taken off the path.  A trajectory at nstep = 10 spent 58 % of its time OUTSIDE the MD steps
// (profiles/r2_microopt_ab.txt (21)): the momenta made five trips through the global workspace (drawn, summed, copied in,
// copied out, summed), the Box-Muller angle, the two actions and the wraps went through library routines with slow paths.
// Here the momenta are drawn into, and stay in, the shared-memory gradient plane; the angle uses sincos_fast, the actions
// cos_fast, the wraps the division-free regularize1_fast (bit-identical).  Same summation orders as hmc_trajectory.
template <class E>
FT_HD void hmc_trajectory_plain(Engine<E>& en, const TrajIO& io) {
    auto& ex = en.ex;
    const int V = en.Vg, L1 = en.L1, LP = en.LP;
    double* X = en.sm(en.oX);
    double* P = en.sm(en.oGR);
    auto action = [&]() {
        double acc = 0.0;
        const int d0 = ex.nt() / L1, d1 = ex.nt() - d0 * L1;
        int n0 = ex.tid() / L1, n1 = ex.tid() - n0 * L1;
        for (int i = ex.tid(); i < V; i += ex.nt(), n0 += d0, n1 += d1) {
            if (n1 >= L1) { n1 -= L1; ++n0; }
            acc += cos_fast(en.plaq(en.oX, n0, n1, 1));
        }
        return -io.beta * ex.sum(acc);
    };
    if (io.first) en.load_field(en.oX, io.field_in);
    ex.sync();
    en.for_links([&](int si, int gi) { en.wsX0[gi] = X[si]; });   // the trajectory's start field, restored on reject
    if (io.p_in) en.for_links([&](int si, int gi) { P[si] = io.p_in[gi]; });
    else {
        Philox ph{ (uint32_t)io.seed, (uint32_t)(io.seed >> 32) };
        for (int j = ex.tid(); j < V; j += ex.nt()) {             // pair j = links 2j, 2j+1 of the (2, L0, L1) layout (same row: L1 is even)
            const int row = (2 * j) / L1, n1 = 2 * j - row * L1;
            uint32_t r[4];
            ph.gen((uint32_t)j, (uint32_t)io.traj, (uint32_t)io.chain, (uint32_t)(io.chain >> 32) ^ 0x5EEDu, r);
            const double u1 = u53(r[0], r[1]), u2 = u53(r[2], r[3]);
            const double rad = sqrt(-2.0 * log(u1));
            double sn, cs;
            sincos_fast(TWO_PI_D * u2, sn, cs);
            P[row * LP + n1] = rad * cs; P[row * LP + n1 + 1] = rad * sn;
        }
    }
    ex.sync();
    double k0 = 0.0;
    en.for_links([&](int si, int) { k0 += P[si] * P[si]; });
    k0 = ex.sum(k0);
    const double s0 = action();
    const double h0 = s0 + 0.5 * k0;
    leapfrog_plain_resident(en, io.beta, io.dt, io.nstep);
    en.for_links([&](int si, int) { X[si] = regularize1_fast(X[si]); });
    ex.sync();
    double k1 = 0.0;
    en.for_links([&](int si, int) { k1 += P[si] * P[si]; });
    k1 = ex.sum(k1);
    const double s1 = action();
    const double h1 = s1 + 0.5 * k1;
    const double u = io.u_in ? *io.u_in : philox_uniform(io);
    const double dH = h1 - h0;
    const double e = exp(-dH);
    const bool acc = u < e;
    if (io.p_out) en.for_links([&](int si, int gi) { io.p_out[gi] = P[si]; });
    if (!acc) {
        ex.sync();
        en.for_links([&](int si, int gi) { X[si] = en.wsX0[gi]; });   // newx = x (bit-identical input)
    }
    ex.sync();
    double qs = 0.0;
    for (int i = ex.tid(); i < V; i += ex.nt()) { int n0, n1; en.site_map(i, n0, n1); qs += regularize1_fast(en.plaq(en.oX, n0, n1, 1)); }
    const double q = floor(0.1 + ex.sum(qs) / TWO_PI_D);
    if (io.last) en.store_field(io.field_out, en.oX);
    if (ex.tid() == 0) {
        if (io.out_dH) *io.out_dH = dH;
        if (io.out_expmdH) *io.out_expmdH = e;
        if (io.out_acc) *io.out_acc = acc ? 1 : 0;
        if (io.out_plaq) *io.out_plaq = (acc ? s1 : s0) / (-io.beta * V);
        if (io.out_Q) *io.out_Q = q;
        if (io.out_h0) *io.out_h0 = h0;
        if (io.out_h1) *io.out_h1 = h1;
    }
    ex.sync();
}

// plain HMC trajectory (hmc_2dU1.py:144-155)
template <class E, bool FAST = false>
FT_HD void hmc_trajectory(Engine<E>& en, const TrajIO& io) {
    auto& ex = en.ex;
    const int V = en.Vg;
    double* P = en.wsP;
    double* X = en.sm(en.oX);
    if (io.first) en.load_field(en.oX, io.field_in);
    ex.sync();
    en.for_links([&](int si, int gi) { en.wsX0[gi] = X[si]; });   // the trajectory's start field, restored on reject
    if (io.p_in) en.for_links([&](int, int gi) { P[gi] = io.p_in[gi]; });
    else philox_momenta(en, io, P);
    ex.sync();
    double k0 = 0.0;
    en.for_links([&](int, int gi) { k0 += P[gi] * P[gi]; });
    k0 = ex.sum(k0);
    double s0 = en.wilson_action(io.beta, 1);
    double h0 = s0 + 0.5 * k0;
    if constexpr (E::kCluster) leapfrog_resident(en, io.dt, io.nstep, P, [&]() { en.wilson_force(io.beta, 1); });
    else leapfrog_plain_fused<E, FAST>(en, io.beta, io.dt, io.nstep, P);
    en.for_links([&](int si, int gi) { X[si] = regularize1(X[si]); });
    ex.sync();
    double k1 = 0.0;
    en.for_links([&](int, int gi) { k1 += P[gi] * P[gi]; });
    k1 = ex.sum(k1);
    double s1 = en.wilson_action(io.beta, 1);
    double h1 = s1 + 0.5 * k1;
    double u = io.u_in ? *io.u_in : philox_uniform(io);
    double dH = h1 - h0;
    double e = exp(-dH);
    bool acc = u < e;
    if (!acc) {
        ex.sync();
        en.for_links([&](int si, int gi) { X[si] = en.wsX0[gi]; });   // newx = x (bit-identical input)
    }
    ex.sync();
    double q = en.topo_floor();
    if (io.last) en.store_field(io.field_out, en.oX);
    if (io.p_out) en.for_links([&](int, int gi) { io.p_out[gi] = P[gi]; });
    if (ex.tid() == 0 && en.rk == 0) {
        if (io.out_dH) *io.out_dH = dH;
        if (io.out_expmdH) *io.out_expmdH = e;
        if (io.out_acc) *io.out_acc = acc ? 1 : 0;
        if (io.out_plaq) *io.out_plaq = (acc ? s1 : s0) / (-io.beta * V);
        if (io.out_Q) *io.out_Q = q;
        if (io.out_h0) *io.out_h0 = h0;
        if (io.out_h1) *io.out_h1 = h1;
    }
    ex.sync();
}

}  // namespace fthmc
