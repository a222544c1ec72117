// expect_math.cuh -- scalar helpers of the Tsit5 step (third part of the NVRTC translation unit,
// after the prelude; also compiled for the HOST by tests/test_device_math.py with g++ -mfma so
// the approximations below are checked against libm on the CPU box).
//
// What they serve: the reference's step-size controller evaluates two `pow` and one `sqrt` per
// attempted step (OrdinaryDiffEq PIController behind src/expectation.jl:191 / :290) and one
// division per state component in the error norm.  FP64 instruction economy decides this path
// (ncu, profiles/): a 64-bit literal costs two moves every time it is used, and CUDA's
// log/exp/div/sqrt each carry a special-case branch.  So
//   (1) every constant of the step lives in __constant__ memory and is read as a c[bank][off]
//       operand of the DFMA itself, or is a value whose low 32 bits are zero (a DFMA immediate);
//   (2) the controller's power function is ONE log and ONE exp, both table-driven: a 128-entry
//       table in shared memory shrinks the argument so far that a degree-5 / degree-4 polynomial
//       reaches 1e-15 (9 + 8 FP64 instructions instead of 23 + 22 for the series without tables);
//   (3) 1/x is MUFU.RCP64H + one cubic Newton step (3 DFMA).
// Absolute error of log_pos and relative error of exp_small are below 4e-15 on the domain the
// controller can reach; a relative perturbation d of a step size moves a trajectory by ~5 d * (its
// global error ~1e-5), and flips an accept/reject or last-step decision with probability ~d per
// step, so 1e-14 keeps both the 1e-8 parity of sums and the bit-exact counters at 1e9 samples.

#ifndef B200E_EXPECT_MATH_CUH
#define B200E_EXPECT_MATH_CUH

#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
#define B200E_MFN __device__ __forceinline__
#define B200E_CONST __constant__
#else
// ---- host twin (tests only) ----
#include <cmath>
#include <cstdint>
#include <cstring>
#define B200E_MFN static inline
#define B200E_CONST static
#define __align__(n) __attribute__((aligned(n)))
static inline int __double2hiint(double x) { int64_t b; std::memcpy(&b, &x, 8); return (int)(b >> 32); }
static inline int __double2loint(double x) { int64_t b; std::memcpy(&b, &x, 8); return (int)(b & 0xffffffff); }
static inline double __hiloint2double(int hi, int lo) {
    const uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double x; std::memcpy(&x, &b, 8); return x;
}
#endif

#define R_(x) ((real)(x))

namespace b200e {

namespace lit {
constexpr double c1 = 0.161, c2 = 0.327, c3 = 0.9, c4 = 0.9800255409045097;
constexpr double a21 = 0.161;
constexpr double a31 = -0.008480655492356989, a32 = 0.335480655492357;
constexpr double a41 = 2.8971530571054935, a42 = -6.359448489975075, a43 = 4.3622954328695815;
constexpr double a51 = 5.325864828439257, a52 = -11.748883564062828, a53 = 7.4955393428898365,
                 a54 = -0.09249506636175525;
constexpr double a61 = 5.86145544294642, a62 = -12.92096931784711, a63 = 8.159367898576159,
                 a64 = -0.071584973281401, a65 = -0.028269050394068383;
constexpr double a71 = 0.09646076681806523, a72 = 0.01, a73 = 0.4798896504144996,
                 a74 = 1.379008574103742, a75 = -3.290069515436081, a76 = 2.324710524099774;
constexpr double bt1 = -0.00178001105222577714, bt2 = -0.0008164344596567469,
                 bt3 = 0.007880878010261995, bt4 = -0.1447110071732629, bt5 = 0.5823571654525552,
                 bt6 = -0.45808210592918697, bt7 = 0.015151515151515152;
// PI controller on e2 = EEst^2:  EEst^b1 = e2^(b1/2)
constexpr double hb1 = 0.5 * 7.0 / 50.0, hb2 = 0.5 * 2.0 / 25.0;
constexpr double gamma = 9.0 / 10.0, qmin = 1.0 / 5.0, qmax = 10.0;
constexpr double ln_gamma = -0.10536051565782630123;  // log(9/10)
constexpr double le2_qoldinit = -18.420680743952364;  // log(1e-4^2)
constexpr double inv_n = 1.0 / B200E_NSTATE;
constexpr double ln2 = 0.6931471805599453094;
constexpr double log2e_128 = 128.0 * 1.4426950408889634074, ln2_128 = 0.6931471805599453094 / 128.0;
constexpr double rnd_magic = 6755399441055744.0;  // 1.5 * 2^52: adds round-to-nearest-integer
constexpr double two52_1023 = 4503599627370496.0 + 1023.0;
constexpr double ln10_04 = 0.4 * 2.302585092994045684;
// The controller never divides the sum of squared scaled residuals by N when N is a power of two: its "e2" is then
// N EEst^2 -- an exact scaling -- and the shift ln N sits in the constants instead: the accept threshold is N (1 + eps),
// the history of a rejected step reads ln N instead of 0, qoldinit and ln(gamma) carry their share.
constexpr int ilog2c(int n) { return n <= 1 ? 0 : 1 + ilog2c(n >> 1); }
constexpr bool e2_scaled = (B200E_NSTATE & (B200E_NSTATE - 1)) == 0;
constexpr double e2_scale = e2_scaled ? (double)B200E_NSTATE : 1.0;
constexpr double ln_n = e2_scaled ? ln2 * ilog2c(B200E_NSTATE) : 0.0;
constexpr double le2_qoldinit_n = le2_qoldinit + ln_n;
constexpr double ln_gamma_n = ln_gamma + (hb1 - hb2) * ln_n;
// series coefficients rounded to the 21 bits a DFMA immediate carries.  Their terms are at most
// 2^-24 (log: r^3/3, |r| <= 2^-8) resp. 2^-27 (exp: r^3/6, |r| <= 2^-8.5) of the result, so the
// 2^-21 rounding stays below 2^-45 * 2^-8 ... i.e. under 1e-14 absolute; they cost no register.
// LambaEM (expect_kernels.cuh): SDE PI-controller exponents on e2 (beta1/2, beta2/2) and the Taylor coefficients of
// sin(pi r / 2) = r (sp0 + r^2 (sp1 + ... + r^2 sp10)), |r| <= 1 (next term 1.3e-18)
constexpr double sde_hb1 = 0.5 * 7.0 / 10.0, sde_hb2 = 0.5 * 2.0 / 5.0;
constexpr double sde_ln_gamma_n = ln_gamma + (sde_hb1 - sde_hb2) * ln_n;
constexpr double sp0 = 1.5707963267948966192, sp1 = -0.64596409750624625366, sp2 = 0.079692626246167045121,
                 sp3 = -0.0046817541353186881007, sp4 = 0.00016044118478735982187, sp5 = -3.5988432352120853405e-6,
                 sp6 = 5.6921729219679268118e-8, sp7 = -6.6880351098114672325e-10, sp8 = 6.0669357311061956671e-12,
                 sp9 = -4.3770654673137422772e-14, sp10 = 2.5714228928604738662e-16;
constexpr double third_i = 0x1.55555p-2, fifth_i = 0x1.9999ap-3;
constexpr double sixth_i = 0x1.55555p-3, r24_i = 0x1.55555p-5;
}  // namespace lit

#if B200E_FP32
#define K_(name) ((real)lit::name)
#else
struct b200e_consts {
    double c1, c2, c3, c4, a21, a31, a32, a41, a42, a43, a51, a52, a53, a54, a61, a62, a63, a64, a65, a71,
        a72, a73, a74, a75, a76, bt1, bt2, bt3, bt4, bt5, bt6, bt7;
    double hb1, hb2, ln_gamma_n, sde_ln_gamma_n, qmin, qmax, le2_qoldinit_n, inv_n;
    double ln2, log2e_128, ln2_128, ln10_04, fifth, r24;
    double sde_hb1, sde_hb2, sp0, sp1, sp2, sp3, sp4, sp5, sp6, sp7, sp8, sp9, sp10;
};
// not `const`: the compiler must read it from the constant bank instead of folding literals back in
B200E_CONST b200e_consts KC = {
    lit::c1, lit::c2, lit::c3, lit::c4, lit::a21, lit::a31, lit::a32, lit::a41, lit::a42, lit::a43, lit::a51,
    lit::a52, lit::a53, lit::a54, lit::a61, lit::a62, lit::a63, lit::a64, lit::a65, lit::a71, lit::a72, lit::a73,
    lit::a74, lit::a75, lit::a76, lit::bt1, lit::bt2, lit::bt3, lit::bt4, lit::bt5, lit::bt6, lit::bt7,
    lit::hb1, lit::hb2, lit::ln_gamma_n, lit::sde_ln_gamma_n, lit::qmin, lit::qmax, lit::le2_qoldinit_n, lit::inv_n,
    lit::ln2, lit::log2e_128, lit::ln2_128, lit::ln10_04, 0.2, 1.0 / 24.0,
    lit::sde_hb1, lit::sde_hb2, lit::sp0, lit::sp1, lit::sp2, lit::sp3, lit::sp4, lit::sp5, lit::sp6, lit::sp7, lit::sp8,
    lit::sp9, lit::sp10};
#define K_(name) (KC.name)
#endif

// ---- range-reduction tables (one copy per block in shared memory, filled at kernel start) ----
constexpr int TAB_BITS = 7, TAB_N = 1 << TAB_BITS;
struct __align__(16) b200e_lgent { double c, lc; };  // c_j = 1/(1 + (j+1/2)/128),  lc_j = -log c_j
struct b200e_tabs {
    b200e_lgent lg[TAB_N];
    double ex[TAB_N];  // 2^(j/128)
};
B200E_MFN void tab_fill(b200e_tabs &T, int j) {
    const double mid = (j + 0.5) / TAB_N;
    T.lg[j].c = 1.0 / (1.0 + mid);
    T.lg[j].lc = log1p(mid);  // = -log(c_j) up to the rounding of the division (1e-16 relative in c)
    T.ex[j] = exp2((double)j / TAB_N);
}

// ---- branch-free scalar helpers -----------------------------------------------------------
#if B200E_FP32
// FP32 mode: the special-function unit's approximations (MUFU.RCP / LG2 / EX2 / RSQ, ~2^-22 relative) are at the
// level of the arithmetic itself and carry no slow-path subroutine; the log clamps its argument so that 0, inf
// and NaN error estimates saturate the controller's clamps exactly as in the FP64 path.
#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
B200E_MFN float rcp_pos(float b) { return __frcp_rn(b); }
B200E_MFN float log_pos(const b200e_tabs &, float x) {
    const float xc = __int_as_float(min(max(__float_as_int(x), 0x00800000), 0x7f000000));  // [2^-126, 2^127]; NaN -> 2^127
    return __logf(xc);
}
B200E_MFN float exp_small(const b200e_tabs &, float y) { return __expf(y); }
B200E_MFN float sqrt_pos(float x) { return __fsqrt_rn(x); }
B200E_MFN float sin_halfpi(float t) { return sinpif(0.5f * t); }
B200E_MFN void sqrt_rsqrt_pos(float x, float &s, float &rs) { s = __fsqrt_rn(x); rs = __frsqrt_rn(x); }
#else
B200E_MFN float rcp_pos(float b) { return 1.0f / b; }
B200E_MFN float log_pos(const b200e_tabs &, float x) { return logf(x); }
B200E_MFN float exp_small(const b200e_tabs &, float y) { return expf(y); }
B200E_MFN float sqrt_pos(float x) { return sqrtf(x); }
B200E_MFN float sin_halfpi(float t) { return sinf(1.5707963267948966f * t); }
B200E_MFN void sqrt_rsqrt_pos(float x, float &s, float &rs) { s = sqrtf(x); rs = 1.0f / s; }
#endif
#else
// 1/b for normal b > 0 to 2^-39.8 (1.1e-12): MUFU.RCP64H seed (20 bits: it reads the high word only) + ONE Newton step,
//   e = 1 - b r;  1/b = r (1 + e + O(e^2)),  e <= 2^-19.9.
// Accuracy matched to the use: rcp_pos, log_pos and exp_small only ever touch the STEP-SIZE CONTROL (scaled error norm,
// PI controller, initial-step heuristic) -- they decide which dt is tried next, never the arithmetic of the solution.
// A relative perturbation d of dt moves a trajectory's end state by ~5 d x (its global error, <= 1e-3 at the default
// tolerance): 1e-12 here is 1e-14 there, six orders below the 1e-8 parity contract and far below the ~1e-6 noise the
// very first step's error estimate carries in any implementation.  Round 1 refined to 1 ulp with
// a cubic step and longer series; the shorter forms are 4 FP64 instructions less per attempted step (118 -> 114 on
// linear2): -2.6 % kernel time on linear2, -2.0 % on lorenz63 (profiles/r2_sweep_mathvar.log), with every counter of the
// GPU suite and of the fuzz runs still equal to the oracle's.
B200E_MFN double rcp_pos(double b) {
    double r;
#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
#else
    r = __hiloint2double(__double2hiint(1.0 / __hiloint2double(__double2hiint(b), 0)), 0);  // same 20-bit class
#endif
    const double e = fma(-b, r, 1.0);
    return fma(r, e, r);
}
// sqrt(x) for normal x > 0: MUFU.RSQ64H seed + two coupled Newton steps (Goldschmidt form)
B200E_MFN double sqrt_pos(double x) {
    double y;
#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#else
    y = __hiloint2double(__double2hiint(1.0 / std::sqrt(__hiloint2double(__double2hiint(x), 0))), 0);
#endif
    // one coupled step takes the 2^-20 seed to 2^-41 in both g ~ sqrt(x) and h ~ 1/(2 sqrt(x)); the correction
    // g + (x - g^2) h then squares that: full double precision (round 1 ran a second coupled step it did not need)
    double g = x * y, h = 0.5 * y;
    const double r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    const double d = fma(-g, g, x);
    return fma(d, h, g);
}
// sqrt(x) and 1/sqrt(x) together (the Goldschmidt iteration above carries both)
B200E_MFN void sqrt_rsqrt_pos(double x, double &s, double &rs) {
    double y;
#if defined(__CUDACC__) || defined(__CUDACC_RTC__)
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#else
    y = __hiloint2double(__double2hiint(1.0 / std::sqrt(__hiloint2double(__double2hiint(x), 0))), 0);
#endif
    double g = x * y, h = 0.5 * y;
    const double r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    const double d = fma(-g, g, x);
    s = fma(d, h, g);   // full precision (see sqrt_pos)
    rs = h + h;         // 2 h -> 1/sqrt(x) to ~1.5e-12: it only scales the LambaEM error estimate (step-size control)
}
// log(x), x > 0:  x = 2^e m, m in [1,2), j = top 7 bits of m's fraction;
//   log x = e ln2 + lc_j + log1p(r),  r = m c_j - 1  (one exact-product DFMA), |r| <= 2^-8,
//   log1p(r) = r - r^2/2 + r^3/3 - r^4/4  (next term 2^-40/5 = 1.8e-13 absolute: step-size control only, see rcp_pos).
// Every bit pattern gives a finite result: zero / subnormal x reads as 2^-1023 m (about -709), +inf and NaN as
// 2^1024 m (about +710), sign-bit patterns higher still -- for all of which the controller's clamps decide alone
// (EEst = 0 -> q = 1/qmax, EEst = Inf -> dt/5 as in the reference).
B200E_MFN double log_pos(const b200e_tabs &T, double x) {
    const unsigned hi = (unsigned)__double2hiint(x);
    const b200e_lgent ent = T.lg[(hi >> (20 - TAB_BITS)) & (TAB_N - 1)];
    // the argument's mantissa under the exponent of 1.0: m in [1, 2) for EVERY bit pattern, so the exponent needs no clamp
    // (round 1 scaled the table's reciprocal by 2^-e instead, which needed one: a VIMNMX and an IADD3 more per call)
    const double m = __hiloint2double((int)((hi & 0x000fffffu) | 0x3ff00000u), __double2loint(x));
    const double r = fma(m, ent.c, -1.0);
    const double ef = __hiloint2double(0x43300000, (int)(hi >> 20)) - lit::two52_1023;  // (double)e, exact
    // Horner: every DFMA has two register sources (accumulator, r) and one immediate -> 2-cycle issue
    double h = fma(r, -0.25, lit::third_i);
    h = fma(h, r, -0.5);
    h = fma(h, r, 1.0);
    return fma(h, r, fma(ef, K_(ln2), ent.lc));
}
// exp(y), |y| < 600:  n = round(128 y / ln2),  exp y = 2^(n >> 7) * 2^((n & 127)/128) * exp(r),
//   r = y - n ln2/128, |r| <= ln2/256,  exp r - 1 = r + r^2 (1/2 + r/6)  (next term r^4/24 <= 2.2e-12 relative)
B200E_MFN double exp_small(const b200e_tabs &T, double y) {
    const double t = fma(y, K_(log2e_128), lit::rnd_magic);
    const int n = __double2loint(t);
    const double nf = t - lit::rnd_magic;
    const double r = fma(-nf, K_(ln2_128), y);
    const double tj = T.ex[n & (TAB_N - 1)];
    double h = fma(r, lit::sixth_i, 0.5);
    h = fma(h, r, 1.0);
    const double p = fma(tj * h, r, tj);
    return __hiloint2double(__double2hiint(p) + ((n >> TAB_BITS) << 20), __double2loint(p));
}
// sin(pi t / 2) for |t| < 2^50, the one transcendental of the KKL noise path (LambaEM evaluates W(t) at arbitrary
// t): n = round(t / 2), r = t - 2 n in [-1, 1] (exact), sin(pi t / 2) = (-1)^n sin(pi r / 2) by an odd Taylor
// polynomial of degree 21 -- 15 FP64 instructions, every coefficient a constant-bank operand, the sign applied on
// the integer pipe.  Absolute error below 3e-16 (tests/test_device_math.py).  CUDA's sincospi costs about twice
// that in FP64 instructions plus two moves per 64-bit literal coefficient.
B200E_MFN double sin_halfpi(double t) {
    const double m = fma(t, 0.5, lit::rnd_magic);
    const int n = __double2loint(m);
    const double r = fma(m - lit::rnd_magic, -2.0, t);
    const double r2 = r * r;
    double h = fma(r2, K_(sp10), K_(sp9));
    h = fma(h, r2, K_(sp8));
    h = fma(h, r2, K_(sp7));
    h = fma(h, r2, K_(sp6));
    h = fma(h, r2, K_(sp5));
    h = fma(h, r2, K_(sp4));
    h = fma(h, r2, K_(sp3));
    h = fma(h, r2, K_(sp2));
    h = fma(h, r2, K_(sp1));
    h = fma(h, r2, K_(sp0));
    const double s = h * r;
    return __hiloint2double(__double2hiint(s) ^ (int)((unsigned)n << 31), __double2loint(s));
}
#endif

// min / max of two NON-NEGATIVE reals through their bit patterns (IEEE order == integer order
// for x >= 0): integer-pipe compares + selects instead of a DSETP on the FP64 pipe plus the
// NaN-fixup sequence fmin/fmax compile to.  A positive-NaN pattern orders above +inf.
// ---- comparisons on the integer pipe --------------------------------------------------------
// A DSETP occupies the FP64 pipe for 2 cycles; integer compares on the bit patterns ride along for
// free (profiles/r1_dp_operands_ubench.txt).  IEEE order equals integer order for non-negative
// values, NaN patterns order above +inf, and for two NEGATIVE values the order is reversed.
#if B200E_FP32
B200E_MFN float abs_bits(float x) { return fabsf(x); }
B200E_MFN float pmin(float a, float b) { return __int_as_float(min(__float_as_int(a), __float_as_int(b))); }
B200E_MFN float pmax(float a, float b) { return __int_as_float(max(__float_as_int(a), __float_as_int(b))); }
B200E_MFN bool same_value(float a, float b) { return a == b; }
B200E_MFN bool le_nonneg(float a, float b) { return a <= b; }
B200E_MFN bool lt_nonneg(float a, float b) { return a < b; }
B200E_MFN bool is_negative(float a) { return a < 0.0f; }
B200E_MFN bool lt_any_pos(float a, float b) { return a < b; }
B200E_MFN float max_with_negative(float a, float q) { return fmaxf(a, q); }
#elif defined(__CUDACC__) || defined(__CUDACC_RTC__)
// |x| through the sign bit (an ALU LOP3; fabs of a double is a DADD on the FP64 pipe)
B200E_MFN double abs_bits(double x) { return __hiloint2double(__double2hiint(x) & 0x7fffffff, __double2loint(x)); }
B200E_MFN double pmin(double a, double b) {
    return __longlong_as_double(min(__double_as_longlong(a), __double_as_longlong(b)));
}
B200E_MFN double pmax(double a, double b) {
    return __longlong_as_double(max(__double_as_longlong(a), __double_as_longlong(b)));
}
// a == b for finite values of equal sign of zero (used for t + dt == t)
B200E_MFN bool same_value(double a, double b) { return __double_as_longlong(a) == __double_as_longlong(b); }
// a <= b / a < b for a >= 0 or NaN (false for every NaN pattern), b >= 0 finite
B200E_MFN bool le_nonneg(double a, double b) {
    return (unsigned long long)__double_as_longlong(a) <= (unsigned long long)__double_as_longlong(b);
}
B200E_MFN bool lt_nonneg(double a, double b) {
    return (unsigned long long)__double_as_longlong(a) < (unsigned long long)__double_as_longlong(b);
}
B200E_MFN bool is_negative(double a) { return __double2hiint(a) < 0; }  // sign bit (true for -0.0)
// a < b for any non-NaN a and a POSITIVE b, as ONE signed 64-bit integer compare of the bit patterns: a pattern with the
// sign bit set is a negative integer (so every negative a, -0.0 included, compares below b), and for a >= 0 the IEEE
// order is the integer order
B200E_MFN bool lt_any_pos(double a, double b) { return __double_as_longlong(a) < __double_as_longlong(b); }
// max(a, q) for finite a and a NEGATIVE constant q: a >= 0 has the smaller unsigned pattern (no sign
// bit), and of two negative values the smaller pattern is the smaller magnitude, i.e. the larger value
B200E_MFN double max_with_negative(double a, double q) {
    return __longlong_as_double((long long)min((unsigned long long)__double_as_longlong(a),
                                               (unsigned long long)__double_as_longlong(q)));
}
#endif

}  // namespace b200e
#endif
