// expect_kernels.cuh -- second part of the NVRTC translation unit, placed AFTER the user's model
// source (rhs / hmap / observable / diffusion are visible here and inline into the loops).
//
// Replaces, for a whole batch per launch, what the reference does per trajectory on a CPU thread:
//   prob_func   src/expectation.jl:175-178 (MC :277-280, process noise :215-225 / :304-314)
//   solve       src/expectation.jl:191, :238, :290, :323  (Tsit5 / Euler-Maruyama)
//   output_func src/expectation.jl:180, :227 (g * pdf)  and :282, :316 (g)
//   reduction   src/expectation.jl:194 (dx .= ...)  and :293, :324 (mean(sol.u) -> sums here)
//
// Execution model: ONE trajectory per thread, state and step control in registers.  Every lane
// owns the static sample sequence i = tid, tid+T, tid+2T, ... (T = grid*block), so the mapping
// sample -> lane -> accumulator is fixed by the launch shape alone: sums are reproducible and
// the step counters are exact integers.  Adaptive step counts differ between lanes; instead of a
// nested per-trajectory loop (warp runs as long as its slowest lane) the loop body is "one RK
// step of my current trajectory", and idle lanes are re-loaded together once at least
// `refill_threshold` lanes of the warp are idle -- the reload (h, pdf, initial-dt heuristic, two
// extra RHS evaluations) costs about one step and would otherwise run under a divergent mask for
// every single finishing lane.

#define B200E_FULL 0xffffffffu

namespace b200e {

constexpr int N = B200E_NSTATE;
constexpr int NP = B200E_NPARAM;
constexpr int NX = B200E_NX;
constexpr int NOUT = B200E_NOUT;
constexpr int NACC = 2 * NOUT;
constexpr int NWARPS = B200E_BLOCK / 32;

__device__ __forceinline__ double neg_inf() { return __longlong_as_double(0xfff0000000000000LL); }

// pdf(d, x) = exp(sum_j logpdf(d_j, x_j))   (src/distribution_utils.jl:42, one exp for the product)
__device__ __forceinline__ double pdf_weight(const b200e_kparams &P, const double *xs) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < NX; j++) {
        const int kind = P.pdf[j].kind;
        if (kind != 0) {
            const double xj = xs[j];
            const bool inside = (xj >= P.pdf[j].lo) && (xj <= P.pdf[j].hi);
            const double z = (xj - P.pdf[j].mu) * P.pdf[j].inv_sigma;
            const double lp = fma(-0.5 * z, z, P.pdf[j].logc);
            s += inside ? lp : neg_inf();
        }
    }
    return exp(s);
}

__device__ __forceinline__ void load_point(const b200e_kparams &P, long long i, double *xs) {
    if (P.layout == 0) {
        const double *xp = P.x + i * NX;
#pragma unroll
        for (int j = 0; j < NX; j++) xs[j] = __ldg(xp + j);
    } else {
#pragma unroll
        for (int j = 0; j < NX; j++) xs[j] = __ldg(P.x + (long long)j * P.ld + i);
    }
}

// Per-lane running sums of the batch, flushed once per launch.  They are touched once per
// TRAJECTORY (not per step), so they live in shared memory, one conflict-free column per thread,
// and stay out of the register budget of the step loop (registers decide how many warps are
// resident, and resident warps are what hides the FP64 dependency chains).
constexpr bool ACC_SMEM = (NACC * 8 + 4 * 8) * B200E_BLOCK <= 32768;
__shared__ double sh_acc[ACC_SMEM ? NACC : 1][B200E_BLOCK];
__shared__ unsigned long long sh_cnt[ACC_SMEM ? 4 : 1][B200E_BLOCK];

struct Accum {
    // register fallback for wide observables (NACC columns of shared memory would not fit)
    double rs[ACC_SMEM ? 1 : NACC];
    unsigned long long rc[ACC_SMEM ? 1 : 4];
    __device__ __forceinline__ double &val(int k) { return ACC_SMEM ? sh_acc[ACC_SMEM ? k : 0][threadIdx.x] : rs[ACC_SMEM ? 0 : k]; }
    __device__ __forceinline__ unsigned long long &cnt(int j) { return ACC_SMEM ? sh_cnt[ACC_SMEM ? j : 0][threadIdx.x] : rc[ACC_SMEM ? 0 : j]; }
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int k = 0; k < NACC; k++) val(k) = 0.0;
#pragma unroll
        for (int j = 0; j < 4; j++) cnt(j) = 0ull;
    }
    __device__ __forceinline__ void count(unsigned nf, unsigned nacc, unsigned nrej, unsigned nfail) {
        cnt(0) += nf; cnt(1) += nacc; cnt(2) += nrej; cnt(3) += nfail;
    }
};

// observable + weight + (store | accumulate): output_func of the reference
template <bool REDUCE>
__device__ __forceinline__ void emit(const b200e_kparams &P, long long i, const real *u, const real *p,
                                     double wgt, Accum &A) {
    real out[NOUT];
    observable(u, p, (real)P.t1, out);
#pragma unroll
    for (int k = 0; k < NOUT; k++) {
        double v = (double)out[k];
        if (P.weight_by_pdf) v *= wgt;
        if (REDUCE) {
            A.val(k) += v;
            A.val(NOUT + k) = fma(v, v, A.val(NOUT + k));
        } else {
            if (P.layout == 0) P.dx[i * NOUT + k] = v;
            else P.dx[(long long)k * P.ld + i] = v;
        }
    }
}

// warp shuffle + shared-memory block reduction; only per-block partials reach HBM.  Counters are
// folded in both modes; the value accumulators only in reduce mode.
template <bool REDUCE>
__device__ __forceinline__ void block_reduce_store(const b200e_kparams &P, Accum &A) {
    __shared__ double sacc[NWARPS][NACC];
    __shared__ unsigned long long scnt[NWARPS][4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long cnt[4];
    double acc[NACC];
#pragma unroll
    for (int k = 0; k < 4; k++) cnt[k] = A.cnt(k);
#pragma unroll
    for (int k = 0; k < NACC; k++) acc[k] = A.val(k);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        if (REDUCE) {
#pragma unroll
            for (int k = 0; k < NACC; k++) acc[k] += __shfl_down_sync(B200E_FULL, acc[k], off);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) cnt[k] += __shfl_down_sync(B200E_FULL, cnt[k], off);
    }
    if (lane == 0) {
        if (REDUCE) {
#pragma unroll
            for (int k = 0; k < NACC; k++) sacc[warp][k] = acc[k];
        }
#pragma unroll
        for (int k = 0; k < 4; k++) scnt[warp][k] = cnt[k];
    }
    __syncthreads();
    if (REDUCE && threadIdx.x < NACC) {
        double t = 0.0;
        for (int w = 0; w < NWARPS; w++) t += sacc[w][threadIdx.x];
        P.partials[(long long)blockIdx.x * NACC + threadIdx.x] = t;
    }
    if (threadIdx.x < 4) {
        const int k = threadIdx.x;
        unsigned long long t = 0ull;
        for (int w = 0; w < NWARPS; w++) t += scnt[w][k];
        P.cpartials[(long long)blockIdx.x * 4 + k] = t;
    }
}

#if B200E_SOLVER == 0
// ------------------------------------------------------------------------------------------
// Adaptive Tsit5 (Tsitouras 2011) with OrdinaryDiffEq's defaults: Hairer-Wanner initial step,
// PI controller beta1=7/50 beta2=2/25 gamma=9/10 qmin=1/5 qmax=10 qoldinit=1e-4, RMS error norm.
//
// FP64 instruction economy (ncu, profiles/r1_*): a 64-bit literal costs two UMOVs every time it is
// used, and CUDA's log/exp/div/sqrt each carry a special-case branch.  So (1) every constant of
// the step lives in __constant__ memory and is read as a c[bank][off] operand of the DFMA itself,
// (2) the controller's power function is one branch-free log and one branch-free exp whose
// coefficients sit in the same table, valid on the clamped domain the controller can reach, and
// (3) the error scale uses MUFU.RCP64H + two Newton steps instead of an IEEE division.
// ------------------------------------------------------------------------------------------
#define R_(x) ((real)(x))

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
// below / above these e2 the clamps decide alone (q = 1/qmax resp. reject factor = qmin):
//   gamma * e2^-hb1 >= qmax  <=>  e2 <= (gamma/qmax)^(1/hb1) = 1.2e-15 ;  take 1e-30 for margin
//   gamma * e2^-hb1 <= qmin  <=>  e2 >= (gamma/qmin)^(1/hb1) = 2.1e9   ;  take 1e12 for margin
constexpr double e2_lo = 1e-30, e2_hi = 1e12;
// ln 2 = ln2_hi + ln2_lo with ln2_hi on 21 bits: n * ln2_hi is exact and ln2_hi is a DFMA immediate
constexpr double ln2_hi = 0x1.62e43p-1, ln2_lo = -0x1.05c610ca86c39p-29;
// high-order series coefficients rounded to the 21 bits a DFMA immediate carries: their terms are
// below 2^-32 of the sum, so the 2^-21 rounding stays below 2^-53 (and they cost no uniform register)
constexpr double lg6i = 0x1.3b13bp-3, lg7i = 0x1.11111p-3, lg8i = 0x1.e1e1ep-4, lg9i = 0x1.af287p-4;
constexpr double lg5i = 0x1.745d1p-3, ex8i = 0x1.a01ap-16;  // term <= 5e-9 of the sum: error <= 3e-15, the controller needs 1e-13
constexpr double ex9i = 0x1.71de4p-19, ex10i = 0x1.27e5p-22, ex11i = 0x1.ae645p-26, ex12i = 0x1.1eed9p-29;
constexpr double ln2 = 0.6931471805599453094, log2e = 1.4426950408889634074;
constexpr double rnd_magic = 6755399441055744.0;  // 1.5 * 2^52: adds round-to-nearest-integer
constexpr double ln10_04 = 0.4 * 2.302585092994045684;
}  // namespace lit

#if B200E_FP32
#define K_(name) ((real)lit::name)
#else
struct b200e_consts {
    double c1, c2, c3, c4, a21, a31, a32, a41, a42, a43, a51, a52, a53, a54, a61, a62, a63, a64, a65, a71,
        a72, a73, a74, a75, a76, bt1, bt2, bt3, bt4, bt5, bt6, bt7;
    double hb1, hb2, ln_gamma, qmin, qmax, le2_qoldinit, inv_n;
    double ln2_hi, ln2_lo, ln2, log2e, rnd_magic, ln10_04;
    double lg[10];  // atanh series 2/(2k+1), k = 0..9
    double ex[13];  // 1/k!, k = 0..12
};
// not `const`: the compiler must read it from the constant bank instead of folding literals back in
__constant__ b200e_consts KC = {
    lit::c1, lit::c2, lit::c3, lit::c4, lit::a21, lit::a31, lit::a32, lit::a41, lit::a42, lit::a43, lit::a51,
    lit::a52, lit::a53, lit::a54, lit::a61, lit::a62, lit::a63, lit::a64, lit::a65, lit::a71, lit::a72, lit::a73,
    lit::a74, lit::a75, lit::a76, lit::bt1, lit::bt2, lit::bt3, lit::bt4, lit::bt5, lit::bt6, lit::bt7,
    lit::hb1, lit::hb2, lit::ln_gamma, lit::qmin, lit::qmax, lit::le2_qoldinit, lit::inv_n,
    lit::ln2_hi, lit::ln2_lo, lit::ln2, lit::log2e, lit::rnd_magic, lit::ln10_04,
    {2.0, 2.0 / 3, 2.0 / 5, 2.0 / 7, 2.0 / 9, 2.0 / 11, 2.0 / 13, 2.0 / 15, 2.0 / 17, 2.0 / 19},
    {1.0, 1.0, 1.0 / 2, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320, 1.0 / 362880,
     1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600}};
#define K_(name) (KC.name)
#endif

// ---- branch-free scalar helpers -----------------------------------------------------------
#if B200E_FP32
__device__ __forceinline__ float rcp_pos(float b) { return 1.0f / b; }
__device__ __forceinline__ float log_pos(float x) { return logf(x); }
__device__ __forceinline__ float exp_small(float y) { return expf(y); }
__device__ __forceinline__ float sqrt_pos(float x) { return sqrtf(x); }
#else
// 1/b for normal b > 0: MUFU.RCP64H seed (~20 bits) + two Newton steps (20 -> 40 -> 80 bits)
__device__ __forceinline__ double rcp_pos(double b) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
    double e = fma(-b, r, 1.0);
    r = fma(r, e, r);
    e = fma(-b, r, 1.0);
    return fma(r, e, r);
}
// sqrt(x) for normal x > 0: MUFU.RSQ64H seed + two coupled Newton steps (Goldschmidt form)
__device__ __forceinline__ double sqrt_pos(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;
    double r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    const double d = fma(-g, g, x);
    return fma(d, h, g);
}
// log(x) for normal x > 0:  x = 2^e m, m in [sqrt(1/2), sqrt(2));  log m = 2 atanh((m-1)/(m+1))
__device__ __forceinline__ double log_pos(double x) {
    int hi = __double2hiint(x), lo = __double2loint(x);
    int e = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;
    const int up = hi >= 0x3ff6a09f;  // m >= ~sqrt(2): halve it
    hi -= up << 20;
    e += up;
    const double m = __hiloint2double(hi, lo);
    const double s = (m - 1.0) * rcp_pos(m + 1.0);
    const double z = s * s;
    // Estrin: depth 5 instead of 10 dependent DFMAs (4 resident warps cannot hide a long chain)
    const double z2 = z * z, z4 = z2 * z2, z8 = z4 * z4;
    const double q01 = fma(K_(lg[1]), z, 2.0), q23 = fma(K_(lg[3]), z, K_(lg[2]));
    const double q45 = fma(lit::lg5i, z, K_(lg[4])), q67 = fma(lit::lg7i, z, lit::lg6i);
    const double q89 = fma(lit::lg9i, z, lit::lg8i);
    const double q03 = fma(q23, z2, q01), q47 = fma(q67, z2, q45);
    const double p = fma(q89, z8, fma(q47, z4, q03));
    return fma((double)e, K_(ln2), s * p);
}
// exp(y) for |y| < 700 (no overflow / subnormal handling)
__device__ __forceinline__ double exp_small(double y) {
    const double t = fma(y, K_(log2e), K_(rnd_magic));
    const int n = __double2loint(t);
    const double nf = t - K_(rnd_magic);
    double r = fma(-nf, lit::ln2_hi, y);
    r = fma(-nf, K_(ln2_lo), r);
    const double r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    const double q01 = r + 1.0, q23 = fma(K_(ex[3]), r, 0.5);
    const double q45 = fma(K_(ex[5]), r, K_(ex[4])), q67 = fma(K_(ex[7]), r, K_(ex[6]));
    const double q89 = fma(lit::ex9i, r, lit::ex8i), qab = fma(lit::ex11i, r, lit::ex10i);
    const double q03 = fma(q23, r2, q01), q47 = fma(q67, r2, q45);
    const double q8c = fma(lit::ex12i, r4, fma(qab, r2, q89));
    const double p = fma(q8c, r8, fma(q47, r4, q03));
    return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}
#endif

enum { LANE_EMPTY = 0, LANE_ACTIVE = 1, LANE_DONE = 2 };

// prob_func + integrator init for one lane: x -> (u0, p) via h, pdf weight, initial dt, FSAL k1
__device__ __forceinline__ void start_trajectory(const b200e_kparams &P, long long i, real *u, real *p,
                                                 real *k1, real &t, real &dt, real &le2old, double &wgt,
                                                 unsigned &nf) {
    double xs[NX];
    load_point(P, i, xs);
    wgt = P.weight_by_pdf ? pdf_weight(P, xs) : 1.0;
    real xr[NX], u0n[N], pn[B200E_DIM(NP)];
#pragma unroll
    for (int j = 0; j < NX; j++) xr[j] = (real)xs[j];
#pragma unroll
    for (int j = 0; j < N; j++) { u0n[j] = (real)P.u0_nom[j]; u[j] = u0n[j]; }
#pragma unroll
    for (int j = 0; j < NP; j++) { pn[j] = (real)P.p_nom[j]; p[j] = pn[j]; }
    hmap(xr, u0n, pn, u, p);

    t = (real)P.t0;
    const real abstol = (real)P.abstol, reltol = (real)P.reltol, dtmax = (real)P.dtmax;
    // Hairer-Wanner initial step (ode_determine_initdt) on squared norms: one sqrt, one log, one exp
    real f0[N], isk[N];
    rhs(f0, u, p, t);
    real s0 = R_(0.0), s1 = R_(0.0);
#pragma unroll
    for (int j = 0; j < N; j++) {
        isk[j] = rcp_pos(fma(fabs(u[j]), reltol, abstol));
        const real a = u[j] * isk[j], b = f0[j] * isk[j];
        s0 = fma(a, a, s0);
        s1 = fma(b, b, s1);
    }
    // d0 = sqrt(s0/N), d1 = sqrt(s1/N):  d0 < 1e-5  <=>  s0 < N * 1e-10
    const real tiny2 = R_(1e-10) * R_(N);
    real dt0 = (s0 < tiny2 || s1 < tiny2) ? R_(1e-6) : R_(0.01) * sqrt_pos(s0 * rcp_pos(s1));
    dt0 = fmin(dt0, dtmax);
#if B200E_FP32
    const real eps10 = R_(10.0) * R_(1.1920928955078125e-07);
#else
    const real eps10 = R_(10.0) * R_(2.220446049250313e-16);
#endif
    if (dt0 < eps10) {
        dt = R_(1e-6);
        nf += 2u;
    } else {
        real u1[N], f1[N];
#pragma unroll
        for (int j = 0; j < N; j++) u1[j] = fma(dt0, f0[j], u[j]);
        rhs(f1, u1, p, t + dt0);
        real s2 = R_(0.0);
#pragma unroll
        for (int j = 0; j < N; j++) {
            const real d = (f1[j] - f0[j]) * isk[j];
            s2 = fma(d, d, s2);
        }
        // dmax^2 = max(d1^2, d2^2) with d2 = sqrt(s2/N)/dt0
        const real idt0 = rcp_pos(dt0);
        const real m2 = fmax(s1, s2 * idt0 * idt0) * R_(lit::inv_n);
        real dt1;
        if (m2 <= R_(1e-30)) dt1 = fmax(R_(1e-6), dt0 * R_(1e-3));
        else dt1 = exp_small(fma(R_(-0.1), log_pos(m2), -K_(ln10_04)));  // 10^(-(2 + log10 dmax)/5)
        dt = fmin(fmin(R_(100.0) * dt0, dt1), dtmax);
        nf += 3u;
    }
    // Tsit5 initialize!: fsalfirst = f(uprev, p, t) (same value as f0; the reference evaluates it
    // again, so it is counted again)
#pragma unroll
    for (int j = 0; j < N; j++) k1[j] = f0[j];
    le2old = K_(le2_qoldinit);
}

// min / max of two NON-NEGATIVE reals through their bit patterns (IEEE order == integer order
// for x >= 0): integer-pipe compares + selects instead of a DSETP on the FP64 pipe plus the
// NaN-fixup sequence fmin/fmax compile to.  A positive-NaN pattern orders above +inf.
#if B200E_FP32
__device__ __forceinline__ float pmin(float a, float b) { return __int_as_float(min(__float_as_int(a), __float_as_int(b))); }
__device__ __forceinline__ float pmax(float a, float b) { return __int_as_float(max(__float_as_int(a), __float_as_int(b))); }
#else
__device__ __forceinline__ double pmin(double a, double b) {
    return __longlong_as_double(min(__double_as_longlong(a), __double_as_longlong(b)));
}
__device__ __forceinline__ double pmax(double a, double b) {
    return __longlong_as_double(max(__double_as_longlong(a), __double_as_longlong(b)));
}
#endif

// One attempted Tsit5 step of the lane's trajectory (perform_step! + stepsize_controller! +
// step_accept/reject_controller! + the loopheader! checks of the reference's integrator loop).
struct Lane {
    real u[N], k1[N], p[B200E_DIM(NP)];
    real t, dt, le2old;
    int state;
    int failed;  // 0 ok, 1 aborted, 2 NaN error estimate seen: abort at the next loop header
    unsigned nf0, nacc, nrej;  // nf = nf0 (initial-step evaluations) + 6 * (nacc + nrej)
};

__device__ __forceinline__ void tsit5_step(Lane &L, const real t1, const real dtmax, const bool clip_dtmax,
                                           const real abstol, const real reltol, const real snap_tol,
                                           const int maxiters, const real one_plus_eps) {
    real *u = L.u, *k1 = L.k1, *p = L.p;
    const real t = L.t;
    // ---- loopheader!: a NaN error estimate in the previous attempt, the iteration cap, or a step
    // below the resolution of t abort the trajectory (dt <= dtmax holds where dt is proposed) ----
    const real dt = pmin(L.dt, t1 - t);  // clip to the end point
    if (L.failed || (int)(L.nacc + L.nrej) >= maxiters || t + dt == t) {
        L.failed = 1;
        L.state = LANE_DONE;
        return;
    }
    // ---- perform_step!: 6 new stages, FSAL ----
    real k2[N], k3[N], k4[N], k5[N], k6[N], k7[N], tmp[N], un[N];
#pragma unroll
    for (int i = 0; i < N; i++) tmp[i] = fma(dt, K_(a21) * k1[i], u[i]);
    rhs(k2, tmp, p, fma(K_(c1), dt, t));
#pragma unroll
    for (int i = 0; i < N; i++) tmp[i] = fma(dt, fma(K_(a32), k2[i], K_(a31) * k1[i]), u[i]);
    rhs(k3, tmp, p, fma(K_(c2), dt, t));
#pragma unroll
    for (int i = 0; i < N; i++)
        tmp[i] = fma(dt, fma(K_(a43), k3[i], fma(K_(a42), k2[i], K_(a41) * k1[i])), u[i]);
    rhs(k4, tmp, p, fma(K_(c3), dt, t));
#pragma unroll
    for (int i = 0; i < N; i++)
        tmp[i] = fma(dt, fma(K_(a54), k4[i], fma(K_(a53), k3[i], fma(K_(a52), k2[i], K_(a51) * k1[i]))), u[i]);
    rhs(k5, tmp, p, fma(K_(c4), dt, t));
#pragma unroll
    for (int i = 0; i < N; i++)
        tmp[i] = fma(dt,
                     fma(K_(a65), k5[i],
                         fma(K_(a64), k4[i], fma(K_(a63), k3[i], fma(K_(a62), k2[i], K_(a61) * k1[i])))),
                     u[i]);
    rhs(k6, tmp, p, t + dt);
#pragma unroll
    for (int i = 0; i < N; i++)
        un[i] = fma(dt,
                    fma(K_(a76), k6[i],
                        fma(K_(a75), k5[i],
                            fma(K_(a74), k4[i], fma(K_(a73), k3[i], fma(K_(a72), k2[i], K_(a71) * k1[i]))))),
                    u[i]);
    rhs(k7, un, p, t + dt);

    // ---- embedded error estimate: calculate_residuals + RMS norm, kept squared ----
    real ss = R_(0.0);
#pragma unroll
    for (int i = 0; i < N; i++) {
        const real ut =
            dt * fma(K_(bt7), k7[i],
                     fma(K_(bt6), k6[i],
                         fma(K_(bt5), k5[i],
                             fma(K_(bt4), k4[i], fma(K_(bt3), k3[i], fma(K_(bt2), k2[i], K_(bt1) * k1[i]))))));
        const real sc = fma(pmax(fabs(u[i]), fabs(un[i])), reltol, abstol);
        const real r = ut * rcp_pos(sc);
        ss = fma(r, r, ss);
    }
    // 1/N is a DFMA immediate when N is a power of two
    const real e2 = ss * (((N & (N - 1)) == 0) ? R_(lit::inv_n) : K_(inv_n));  // EEst^2
    // NaN estimate: the reference counts a rejection, dt turns NaN and the solve aborts at the next
    // loop header -- same here through L.failed = 2 (the accept test below is false for a NaN)
    L.failed = (e2 == e2) ? 0 : 2;
    const bool accept = e2 <= one_plus_eps;  // sqrt(e2) <= 1  <=>  e2 <= nextafter(1)

    // ---- PI controller in log space on e2: one log + one exp instead of two pow + sqrt ----
    // accept:  1/q = gamma * qold^b2 / EEst^b1  clamped to [qmin, qmax]
    // reject:  dt *= max(qmin, gamma / EEst^b1)  (< 1, so the same clamp applies)
    // No special cases: log_pos maps e2 = 0 / subnormal to about -709 and e2 = +inf to +710, for
    // which the clamps on 1/q decide alone exactly as the reference's (EEst == 0 -> q = 1/qmax;
    // EEst = Inf -> dt/5), and |b2*le2old - b1*le2| <= 50 stays far inside exp_small's range.
    const real le2 = log_pos(e2);
    const real hist = accept ? L.le2old : R_(0.0);
    real qinv = exp_small(fma(K_(hb2), hist, fma(-K_(hb1), le2, K_(ln_gamma))));  // gamma * e2old^hb2 / e2^hb1
    qinv = pmin(R_(lit::qmax), pmax(K_(qmin), qinv));
    L.dt = dt * qinv;
    if (clip_dtmax) L.dt = pmin(dtmax, L.dt);  // warp-uniform; only when the caller set dtmax < t1 - t0
    if (accept) {
        L.le2old = fmax(le2, K_(le2_qoldinit));  // qold = max(EEst, 1e-4)
        real tn = t + dt;
        if (fabs(tn - t1) < snap_tol) tn = t1;  // fixed_t_for_floatingpoint_error!
        L.t = tn;
#pragma unroll
        for (int i = 0; i < N; i++) { u[i] = un[i]; k1[i] = k7[i]; }
        L.nacc++;
        if (!(tn < t1)) L.state = LANE_DONE;
    } else {
        L.nrej++;
    }
}

template <bool REDUCE>
__device__ __forceinline__ void tsit5_run(const b200e_kparams &P) {
    long long next = (long long)blockIdx.x * B200E_BLOCK + threadIdx.x;
    long long cur = -1;

    Lane L;
    L.t = R_(0.0); L.dt = R_(0.0); L.le2old = R_(0.0);
    L.state = LANE_EMPTY; L.failed = 0;
    L.nf0 = 0; L.nacc = 0; L.nrej = 0;
    double wgt = 1.0;
    Accum A;
    A.init();

    const real t1 = (real)P.t1, dtmax = (real)P.dtmax;
    const real abstol = (real)P.abstol, reltol = (real)P.reltol;
    const real snap_tol = (real)P.snap_tol;
    const int maxiters = (int)P.maxiters;
    // the clip to the end point (dt <= t1 - t) already enforces the default dtmax = t1 - t0
    const bool clip_dtmax = P.dtmax < P.t1 - P.t0;
#if B200E_FP32
    const real one_plus_eps = R_(1.0) + R_(1.1920928955078125e-07);
#else
    const real one_plus_eps = R_(1.0) + R_(2.220446049250313e-16);
#endif

    // refill policy: an explicit threshold (1..32), or 0 = automatic: the first generation of 32
    // trajectories runs to its slowest lane; if that left the warp idle for more than a fifth of the
    // lane-steps (max > 1.25 mean attempted steps) the warp re-loads whenever 8 lanes are idle from
    // then on, otherwise it keeps starting 32 trajectories together.  Sample -> lane ownership does
    // not depend on the policy, so results and counters are identical either way.
    int thr_eff = P.refill_threshold > 0 ? P.refill_threshold : 32;
    bool tune = P.refill_threshold <= 0;

    for (;;) {
        // ---- refill phase: runs only when the set of stepping lanes has changed ----
        const bool wants = (L.state == LANE_DONE) || (L.state == LANE_EMPTY && next < P.npts);
        const unsigned idle_m = __ballot_sync(B200E_FULL, wants);
        unsigned act_m = __ballot_sync(B200E_FULL, L.state == LANE_ACTIVE);
        if ((idle_m | act_m) == 0u) break;
        const int live = __popc(idle_m | act_m);
        const int thr = thr_eff < live ? thr_eff : live;
        if (__popc(idle_m) >= thr) {  // warp-uniform
            const unsigned done_m = __ballot_sync(B200E_FULL, L.state == LANE_DONE);
            if (tune && done_m == B200E_FULL) {  // first full generation finished: measure its imbalance
                const unsigned st = L.nacc + L.nrej;
                const unsigned mx = __reduce_max_sync(B200E_FULL, st), sm = __reduce_add_sync(B200E_FULL, st);
                if (mx * 128u > sm * 5u) thr_eff = 8;  // max > 1.25 * mean
                tune = false;
            }
            if (wants) {
                if (L.state == LANE_DONE) {
                    emit<REDUCE>(P, cur, L.u, L.p, wgt, A);
                    if (!REDUCE && P.steps) P.steps[cur] = (int)(L.nacc + L.nrej);
                    A.count(L.nf0 + 6u * (L.nacc + L.nrej), L.nacc, L.nrej, L.failed ? 1u : 0u);
                    L.state = LANE_EMPTY;
                }
                if (next < P.npts) {
                    cur = next;
                    next += (long long)gridDim.x * B200E_BLOCK;
                    L.nf0 = 0; L.nacc = 0; L.nrej = 0; L.failed = 0;
                    start_trajectory(P, cur, L.u, L.p, L.k1, L.t, L.dt, L.le2old, wgt, L.nf0);
                    L.state = (L.t < t1) ? LANE_ACTIVE : LANE_DONE;
                }
            }
            act_m = __ballot_sync(B200E_FULL, L.state == LANE_ACTIVE);
        }
        if (act_m == 0u) continue;
        // ---- step phase: a tight loop, one vote per step, until some lane finishes ----
        do {
            if (L.state == LANE_ACTIVE) tsit5_step(L, t1, dtmax, clip_dtmax, abstol, reltol, snap_tol, maxiters, one_plus_eps);
        } while (__ballot_sync(B200E_FULL, L.state == LANE_ACTIVE) == act_m);
    }
    block_reduce_store<REDUCE>(P, A);
}

extern "C" __global__ void __launch_bounds__(B200E_BLOCK, B200E_MINBLOCKS)
b200e_nodes(const __grid_constant__ b200e_kparams P) {
    tsit5_run<false>(P);
}
extern "C" __global__ void __launch_bounds__(B200E_BLOCK, B200E_MINBLOCKS)
b200e_reduce(const __grid_constant__ b200e_kparams P) {
    tsit5_run<true>(P);
}

#else
// ------------------------------------------------------------------------------------------
// Fixed-step Euler-Maruyama driven by the KKL sine series (process-noise path):
//   u += f(u,p,t) dt + g(u,p,t) dW,  dW = W(t+dt) - W(t),  W(t) = sum_k Z_k phi_k(t)
// phi_k(t_n) does not depend on the sample: the host tabulates it once per handle.
// Every lane takes the same number of steps, so the warp never diverges.
// ------------------------------------------------------------------------------------------
constexpr int NK = B200E_NKKL;

template <bool REDUCE>
__device__ __forceinline__ void em_run(const b200e_kparams &P) {
    const long long T = (long long)gridDim.x * B200E_BLOCK;
    Accum A;
    A.init();
    const long long nsteps = P.em_nsteps;
    const real t0 = (real)P.t0, t1 = (real)P.t1, h = (real)P.dt_fixed;
    for (long long i = (long long)blockIdx.x * B200E_BLOCK + threadIdx.x; i < P.npts; i += T) {
        double xs[NX];
        load_point(P, i, xs);
        const double wgt = P.weight_by_pdf ? pdf_weight(P, xs) : 1.0;
        real xr[NX], u0n[N], pn[B200E_DIM(NP)], u[N], p[B200E_DIM(NP)], z[NK];
#pragma unroll
        for (int j = 0; j < NX; j++) xr[j] = (real)xs[j];
#pragma unroll
        for (int j = 0; j < N; j++) { u0n[j] = (real)P.u0_nom[j]; u[j] = u0n[j]; }
#pragma unroll
        for (int j = 0; j < NP; j++) { pn[j] = (real)P.p_nom[j]; p[j] = pn[j]; }
#pragma unroll
        for (int k = 0; k < NK; k++) z[k] = (real)0;
        // Z, p = h(x, u0, p); u0 stays nominal (expectation.jl:216)
        hmap(xr, u0n, pn, z, p);

        const double *tab = P.kkl_table;
        real wcur = (real)0;
#pragma unroll
        for (int k = 0; k < NK; k++) wcur = fma(z[k], (real)__ldg(tab + k), wcur);
        for (long long s = 0; s < nsteps; s++) {
            const real t = fma((real)s, h, t0);
            const real tn = (s + 1 == nsteps) ? t1 : fma((real)(s + 1), h, t0);
            const real dt = tn - t;
            const double *row = tab + (s + 1) * NK;
            real wn = (real)0;
#pragma unroll
            for (int k = 0; k < NK; k++) wn = fma(z[k], (real)__ldg(row + k), wn);
            const real dW = wn - wcur;
            wcur = wn;
            real f[N], g[N];
            rhs(f, u, p, t);
            diffusion(g, u, p, t);
#pragma unroll
            for (int j = 0; j < N; j++) u[j] = fma(g[j], dW, fma(dt, f[j], u[j]));
        }
        int failed = 0;
#pragma unroll
        for (int j = 0; j < N; j++) failed |= !(u[j] == u[j]);
        emit<REDUCE>(P, i, u, p, wgt, A);
        if (!REDUCE && P.steps) P.steps[i] = (int)nsteps;
        A.count((unsigned)nsteps, (unsigned)nsteps, 0u, (unsigned)failed);
    }
    block_reduce_store<REDUCE>(P, A);
}

extern "C" __global__ void __launch_bounds__(B200E_BLOCK, B200E_MINBLOCKS)
b200e_nodes(const __grid_constant__ b200e_kparams P) {
    em_run<false>(P);
}
extern "C" __global__ void __launch_bounds__(B200E_BLOCK, B200E_MINBLOCKS)
b200e_reduce(const __grid_constant__ b200e_kparams P) {
    em_run<true>(P);
}
#endif

}  // namespace b200e
