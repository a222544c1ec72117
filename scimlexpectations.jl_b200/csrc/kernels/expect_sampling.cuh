// expect_sampling.cuh -- counter-based sampling of the product density d, BIT-IDENTICAL on host and
// device.
//
// Replaces: `rand(d)` inside prob_func of solve(exprob, MonteCarlo(n))   (src/expectation.jl:277-280,
// :304-305; src/distribution_utils.jl:43: one independent draw per factor of GenericDistribution),
// hoisted out of the per-trajectory closure: sample i of stream `seed` is a pure function of
// (seed, i), so the kernel draws it in registers (no 8*nx bytes per sample over PCIe / HBM) and the
// host twin b200e_sample_host reproduces the very same doubles for the CPU arm ("the same sample
// sets", north_star) -- Julia's task-local Xoshiro stream could not be reproduced on a GPU anyway.
//
// This one text is compiled three ways: into the NVRTC translation unit of every model (device), into
// libb200expect.so's static sampling kernel (device, nvcc) and into its host code (g++ via nvcc,
// -ffp-contract=off).  Bit-identity rests on using only operations that IEEE 754 defines exactly --
// integer arithmetic, +, *, /, sqrt, fma, each spelled out so that neither compiler may contract or
// reorder them (device: the __d*_rn intrinsics are never fused; host: contraction is switched off) --
// and NO library function: log is restated below with those operations.
//
//   generator  Philox4x32-10 (Salmon et al., SC'11): key = seed, counter = (i_lo, i_hi, j/2, 0);
//              one call yields two 52-bit uniforms u = (2k+1) 2^-53 in (0,1) for dimensions j, j+1
//   Uniform(a,b)            x = fma(b-a, u, a)
//   Normal / truncated      x = mu + sigma * Phi^-1(cdf(lo) + u (cdf(hi) - cdf(lo))), clamped to [lo,hi];
//                           Phi^-1 = Wichura's AS241 PPND16 (relative error 1e-16), cdf(lo), cdf(hi) from the host

#ifndef B200E_EXPECT_SAMPLING_CUH
#define B200E_EXPECT_SAMPLING_CUH

#if defined(__CUDACC_RTC__)
#define B200E_SFN __device__ __forceinline__
#elif defined(__CUDACC__)
#define B200E_SFN __host__ __device__ inline
#else
#define B200E_SFN static inline
#endif

#if defined(__CUDA_ARCH__) || defined(__CUDACC_RTC__)
#define S_MUL(a, b) __dmul_rn((a), (b))
#define S_ADD(a, b) __dadd_rn((a), (b))
#define S_SUB(a, b) __dsub_rn((a), (b))
#define S_FMA(a, b, c) __fma_rn((a), (b), (c))
#define S_DIV(a, b) __ddiv_rn((a), (b))
#define S_SQRT(a) __dsqrt_rn((a))
#define S_MULHI(a, b) __umulhi((a), (b))
#define S_BITS(x) ((unsigned long long)__double_as_longlong(x))
#define S_FROM_BITS(b) __longlong_as_double((long long)(b))
#else
#include <cmath>
#include <cstring>
#define S_MUL(a, b) ((a) * (b))
#define S_ADD(a, b) ((a) + (b))
#define S_SUB(a, b) ((a) - (b))
#define S_FMA(a, b, c) std::fma((a), (b), (c))
#define S_DIV(a, b) ((a) / (b))
#define S_SQRT(a) std::sqrt((a))
#define S_MULHI(a, b) ((unsigned int)(((unsigned long long)(a) * (unsigned long long)(b)) >> 32))
static inline unsigned long long b200e_s_bits(double x) { unsigned long long b; std::memcpy(&b, &x, 8); return b; }
static inline double b200e_s_from_bits(unsigned long long b) { double x; std::memcpy(&x, &b, 8); return x; }
#define S_BITS(x) b200e_s_bits(x)
#define S_FROM_BITS(b) b200e_s_from_bits(b)
#endif

// The coefficients of det_log and of AS241 in ONE list.  On the device they are read as constant-bank operands of
// the FMAs themselves (a 64-bit literal costs two moves every time it is used: 100+ instructions per sample when
// written inline); the host reads a plain array.  Same values, same operations: the twins stay bit-identical.
//   [0..9] log series 2/21 .. 2/3;  [10..11] ln2 hi / lo;  [12..27] central region a7..a0, b7..b0 (b0 = 1);
//   [28..43] intermediate region c7..c0, d7..d0;  [44..59] tail region e7..e0, f7..f0
#define B200E_SMP_COEFFS                                                                                               \
    2.0 / 21.0, 2.0 / 19.0, 2.0 / 17.0, 2.0 / 15.0, 2.0 / 13.0, 2.0 / 11.0, 2.0 / 9.0, 2.0 / 7.0, 2.0 / 5.0, 2.0 / 3.0,   \
    0x1.62e42feep-1, 0x1.a39ef35793c76p-33,                                                                            \
    2509.0809287301226727, 33430.575583588128105, 67265.770927008700853, 45921.953931549871457,                        \
    13731.693765509461125, 1971.5909503065514427, 133.14166789178437745, 3.387132872796366608,                         \
    5226.495278852545925, 28729.085735721942674, 39307.89580009271061, 21213.794301586595867,                          \
    5394.1960214247511077, 687.1870074920579083, 42.313330701600911252, 1.0,                                           \
    7.7454501427834140764e-4, 0.0227238449892691845833, 0.24178072517745061177, 1.27045825245236838258,                \
    3.64784832476320460504, 5.7694972214606914055, 4.6303378461565452959, 1.42343711074968357734,                      \
    1.05075007164441684324e-9, 5.475938084995344946e-4, 0.0151986665636164571966, 0.14810397642748007459,              \
    0.68976733498510000455, 1.6763848301838038494, 2.05319162663775882187, 1.0,                                        \
    2.01033439929228813265e-7, 2.71155556874348757815e-5, 0.0012426609473880784386, 0.026532189526576123093,           \
    0.29656057182850489123, 1.7848265399172913358, 5.4637849111641143699, 6.6579046435011037772,                       \
    2.04426310338993978564e-15, 1.4215117583164458887e-7, 1.8463183175100546818e-5, 7.868691311456132591e-4,           \
    0.0148753612908506148525, 0.13692988092273580531, 0.59983220655588793769, 1.0
#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
__constant__ double b200e_smp_coef_dev[60] = {B200E_SMP_COEFFS};  // not const: read from the bank, not folded back into literals
#endif
#if !defined(__CUDACC_RTC__)
static const double b200e_smp_coef_host[60] = {B200E_SMP_COEFFS};
#endif
#if defined(__CUDA_ARCH__) || defined(__CUDACC_RTC__)
#define S_K(i) b200e_smp_coef_dev[i]
#else
#define S_K(i) b200e_smp_coef_host[i]
#endif

namespace b200e_smp {
// p(r) = ((k[0] r + k[1]) r + ...) r + k[7]: eight coefficients starting at S_K(first)
B200E_SFN double horner8(int first, double r) {
    double a = S_K(first);
#pragma unroll
    for (int i = 1; i < 8; i++) a = S_FMA(a, r, S_K(first + i));
    return a;
}

// ---- Philox4x32-10 ----------------------------------------------------------------------------
B200E_SFN void philox4x32_10(unsigned int c0, unsigned int c1, unsigned int c2, unsigned int c3, unsigned int k0,
                             unsigned int k1, unsigned int out[4]) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const unsigned int hi0 = S_MULHI(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const unsigned int hi1 = S_MULHI(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const unsigned int n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 64 random bits -> u = (2k+1) 2^-53 with k their top 52 bits: strictly inside (0,1), exact in FP64
B200E_SFN double uniform52(unsigned int lo, unsigned int hi) {
    const unsigned long long k = (((unsigned long long)hi << 32) | lo) >> 12;
    return S_MUL((double)(2ull * k + 1ull), 0x1p-53);
}

// ---- log(x) for normal x > 0, from exactly-rounded operations only -------------------------------
//   x = 2^e m, m in [sqrt(1/2), sqrt(2)); s = (m-1)/(m+1); log m = 2 s (1 + z/3 + z^2/5 + ... + z^10/21), z = s^2
//   (|s| <= 0.1716: the first dropped term is below 6e-19).  Error ~1e-16 relative; identical everywhere.
B200E_SFN double det_log(double x) {
    unsigned long long b = S_BITS(x);
    int e = (int)((b >> 52) & 0x7ffull) - 1023;
    b = (b & 0x000fffffffffffffull) | 0x3ff0000000000000ull;
    if ((b >> 32) >= 0x3ff6a09full) { b -= 0x0010000000000000ull; e += 1; }  // m >= ~sqrt(2): halve it
    const double m = S_FROM_BITS(b);
    const double s = S_DIV(S_SUB(m, 1.0), S_ADD(m, 1.0));
    const double z = S_MUL(s, s);
    double p = S_K(0);
#pragma unroll
    for (int i = 1; i < 10; i++) p = S_FMA(p, z, S_K(i));
    const double logm = S_FMA(S_MUL(s, z), p, S_MUL(2.0, s));
    const double ed = (double)e;
    // ln 2 = hi + lo with hi on 32 bits: e * hi is exact
    return S_FMA(ed, S_K(10), S_FMA(ed, S_K(11), logm));
}

// ---- Phi^-1(p), 0 < p < 1: Wichura, Algorithm AS241 (Appl. Statist. 37, 1988), routine PPND16 ------
B200E_SFN double norm_ppf(double p) {
    const double q = S_SUB(p, 0.5);
    const double aq = q < 0.0 ? -q : q;
    if (aq <= 0.425) {
        const double r = S_SUB(0.180625, S_MUL(q, q));
        const double a = horner8(12, r), d = horner8(20, r);
        return S_DIV(S_MUL(q, a), d);
    }
    double r = q < 0.0 ? p : S_SUB(1.0, p);
    r = S_SQRT(-det_log(r));
    double val;
    if (r <= 5.0) {
        r = S_SUB(r, 1.6);
        const double a = horner8(28, r), d = horner8(36, r);
        val = S_DIV(a, d);
    } else {
        r = S_SUB(r, 5.0);
        const double a = horner8(44, r), d = horner8(52, r);
        val = S_DIV(a, d);
    }
    return q < 0.0 ? -val : val;
}

// one factor of the product density (b200e_kpdf: kind 1 uniform, 2 normal, 3 truncated normal)
//   smp_a, smp_b: uniform -> (lo, hi-lo);  normal -> (mu, sigma);  smp_c, smp_d: (cdf(lo), cdf(hi)-cdf(lo))
B200E_SFN double draw(int kind, double lo, double hi, double smp_a, double smp_b, double smp_c, double smp_d, double u) {
    double x;
    if (kind == 1) {
        x = S_FMA(smp_b, u, smp_a);
    } else {
        const double pu = S_FMA(u, smp_d, smp_c);
        x = S_FMA(smp_b, norm_ppf(pu), smp_a);
    }
    x = x < lo ? lo : x;
    return x > hi ? hi : x;
}

// sample `index` of stream `seed`: xs[0..NX) with a compile-time dimension (fully unrolled: xs stays in
// registers inside the model kernels)
template <int NX_, class PDF>
B200E_SFN void sample_point_fixed(const PDF *pdf, unsigned long long seed, unsigned long long index, double *xs) {
#pragma unroll
    for (int j = 0; j < NX_; j += 2) {
        unsigned int w[4];
        philox4x32_10((unsigned int)index, (unsigned int)(index >> 32), (unsigned int)(j >> 1), 0u, (unsigned int)seed,
                      (unsigned int)(seed >> 32), w);
        xs[j] = draw(pdf[j].kind, pdf[j].lo, pdf[j].hi, pdf[j].smp_a, pdf[j].smp_b, pdf[j].smp_c, pdf[j].smp_d,
                     uniform52(w[0], w[1]));
        if (j + 1 < NX_)
            xs[j + 1] = draw(pdf[j + 1].kind, pdf[j + 1].lo, pdf[j + 1].hi, pdf[j + 1].smp_a, pdf[j + 1].smp_b,
                             pdf[j + 1].smp_c, pdf[j + 1].smp_d, uniform52(w[2], w[3]));
    }
}

// the same with a run-time dimension (host twin, static sampling kernel)
template <class PDF>
B200E_SFN void sample_point(const PDF *pdf, int nx, unsigned long long seed, unsigned long long index, double *xs) {
    for (int j = 0; j < nx; j += 2) {
        unsigned int w[4];
        philox4x32_10((unsigned int)index, (unsigned int)(index >> 32), (unsigned int)(j >> 1), 0u, (unsigned int)seed,
                      (unsigned int)(seed >> 32), w);
        xs[j] = draw(pdf[j].kind, pdf[j].lo, pdf[j].hi, pdf[j].smp_a, pdf[j].smp_b, pdf[j].smp_c, pdf[j].smp_d,
                     uniform52(w[0], w[1]));
        if (j + 1 < nx)
            xs[j + 1] = draw(pdf[j + 1].kind, pdf[j + 1].lo, pdf[j + 1].hi, pdf[j + 1].smp_a, pdf[j + 1].smp_b,
                             pdf[j + 1].smp_c, pdf[j + 1].smp_d, uniform52(w[2], w[3]));
    }
}

}  // namespace b200e_smp
#endif
