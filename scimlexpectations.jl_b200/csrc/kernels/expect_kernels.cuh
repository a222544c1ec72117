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

// per-lane running sums of the batch; flushed once per launch
struct Accum {
    double s[NOUT];
    double s2[NOUT];
    unsigned long long nf, nacc, nrej, nfail;
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int k = 0; k < NOUT; k++) { s[k] = 0.0; s2[k] = 0.0; }
        nf = nacc = nrej = nfail = 0ull;
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
            A.s[k] += v;
            A.s2[k] = fma(v, v, A.s2[k]);
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
    unsigned long long cnt[4] = {A.nf, A.nacc, A.nrej, A.nfail};
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        if (REDUCE) {
#pragma unroll
            for (int k = 0; k < NOUT; k++) {
                A.s[k] += __shfl_down_sync(B200E_FULL, A.s[k], off);
                A.s2[k] += __shfl_down_sync(B200E_FULL, A.s2[k], off);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) cnt[k] += __shfl_down_sync(B200E_FULL, cnt[k], off);
    }
    if (lane == 0) {
        if (REDUCE) {
#pragma unroll
            for (int k = 0; k < NOUT; k++) { sacc[warp][k] = A.s[k]; sacc[warp][NOUT + k] = A.s2[k]; }
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
// ------------------------------------------------------------------------------------------
#define R_(x) ((real)(x))
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

constexpr double BETA1 = 7.0 / 50.0, BETA2 = 2.0 / 25.0, GAMMA = 9.0 / 10.0;
constexpr double QMIN = 1.0 / 5.0, QMAX = 10.0;
// log(qoldinit) = log(1e-4)
constexpr double LOG_QOLDINIT = -9.210340371976182;

enum { LANE_EMPTY = 0, LANE_ACTIVE = 1, LANE_DONE = 2 };

// RMS norm of v[i]/sk[i] over the state (ODE_DEFAULT_NORM)
__device__ __forceinline__ real rms_scaled(const real *v, const real *sk) {
    real s = R_(0.0);
#pragma unroll
    for (int i = 0; i < N; i++) {
        const real r = v[i] / sk[i];
        s = fma(r, r, s);
    }
    return sqrt(s / R_(N));
}

// prob_func + integrator init for one lane: x -> (u0, p) via h, pdf weight, initial dt, FSAL k1
__device__ __forceinline__ void start_trajectory(const b200e_kparams &P, long long i, real *u, real *p,
                                                 real *k1, real &t, real &dt, real &lqold, double &wgt,
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
    real f0[N], sk[N];
    rhs(f0, u, p, t);
#pragma unroll
    for (int j = 0; j < N; j++) sk[j] = fma(fabs(u[j]), reltol, abstol);
    const real d0 = rms_scaled(u, sk);
    const real d1 = rms_scaled(f0, sk);
    real dt0 = (d0 < R_(1e-5) || d1 < R_(1e-5)) ? R_(1e-6) : (d0 / d1) / R_(100.0);
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
        real u1[N], f1[N], df[N];
#pragma unroll
        for (int j = 0; j < N; j++) u1[j] = fma(dt0, f0[j], u[j]);
        rhs(f1, u1, p, t + dt0);
#pragma unroll
        for (int j = 0; j < N; j++) df[j] = f1[j] - f0[j];
        const real d2 = rms_scaled(df, sk) / dt0;
        const real dmax = fmax(d1, d2);
        real dt1;
        if (dmax <= R_(1e-15)) dt1 = fmax(R_(1e-6), dt0 * R_(1e-3));
        else dt1 = exp10(-(R_(2.0) + log10(dmax)) / R_(5.0));
        dt = fmin(fmin(R_(100.0) * dt0, dt1), dtmax);
        nf += 3u;
    }
    // Tsit5 initialize!: fsalfirst = f(uprev, p, t) (same value as f0; the reference evaluates it
    // again, so it is counted again)
#pragma unroll
    for (int j = 0; j < N; j++) k1[j] = f0[j];
    lqold = R_(LOG_QOLDINIT);
}

template <bool REDUCE>
__device__ __forceinline__ void tsit5_run(const b200e_kparams &P) {
    const long long T = (long long)gridDim.x * B200E_BLOCK;
    long long next = (long long)blockIdx.x * B200E_BLOCK + threadIdx.x;
    long long cur = -1;

    real u[N], k1[N], p[B200E_DIM(NP)];
    real t = R_(0.0), dt = R_(0.0), lqold = R_(0.0);
    double wgt = 1.0;
    int state = LANE_EMPTY;
    int failed = 0;
    long long iter = 0;
    unsigned tnf = 0, tacc = 0, trej = 0;  // counters of the current trajectory
    Accum A;
    A.zero();

    const real t1 = (real)P.t1, dtmax = (real)P.dtmax;
    const real abstol = (real)P.abstol, reltol = (real)P.reltol;
    const real snap_tol = (real)P.snap_tol;
#if B200E_FP32
    const real one_plus_eps = R_(1.0) + R_(1.1920928955078125e-07);
#else
    const real one_plus_eps = R_(1.0) + R_(2.220446049250313e-16);
#endif

    for (;;) {
        const bool wants = (state == LANE_DONE) || (state == LANE_EMPTY && next < P.npts);
        const unsigned idle_m = __ballot_sync(B200E_FULL, wants);
        const unsigned act_m = __ballot_sync(B200E_FULL, state == LANE_ACTIVE);
        if ((idle_m | act_m) == 0u) break;
        const int live = __popc(idle_m | act_m);
        const int thr = P.refill_threshold < live ? P.refill_threshold : live;
        if (__popc(idle_m) >= thr) {  // warp-uniform
            if (wants) {
                if (state == LANE_DONE) {
                    emit<REDUCE>(P, cur, u, p, wgt, A);
                    if (!REDUCE && P.steps) P.steps[cur] = (int)(tacc + trej);
                    A.nf += tnf; A.nacc += tacc; A.nrej += trej; A.nfail += (unsigned)failed;
                    state = LANE_EMPTY;
                }
                if (next < P.npts) {
                    cur = next;
                    next += T;
                    tnf = 0; tacc = 0; trej = 0; failed = 0; iter = 0;
                    start_trajectory(P, cur, u, p, k1, t, dt, lqold, wgt, tnf);
                    state = (t < t1) ? LANE_ACTIVE : LANE_DONE;
                }
            }
        }
        if (state == LANE_ACTIVE) {
            // ---- loopheader!: iteration cap, dt bounds, clip to the end point ----
            iter++;
            dt = fmin(dt, dtmax);
            dt = fmin(dt, t1 - t);
            const bool bad_dt = !(fabs(dt) <= R_(1.79e308)) || (t + dt == t);  // NaN/Inf/underflow
            if (iter > P.maxiters || bad_dt) {
                failed = 1;
                state = LANE_DONE;
                continue;
            }
            // ---- perform_step!: 6 new stages, FSAL ----
            real k2[N], k3[N], k4[N], k5[N], k6[N], k7[N], tmp[N], un[N];
#pragma unroll
            for (int i = 0; i < N; i++) tmp[i] = fma(dt, R_(a21) * k1[i], u[i]);
            rhs(k2, tmp, p, fma(R_(c1), dt, t));
#pragma unroll
            for (int i = 0; i < N; i++) tmp[i] = fma(dt, fma(R_(a32), k2[i], R_(a31) * k1[i]), u[i]);
            rhs(k3, tmp, p, fma(R_(c2), dt, t));
#pragma unroll
            for (int i = 0; i < N; i++)
                tmp[i] = fma(dt, fma(R_(a43), k3[i], fma(R_(a42), k2[i], R_(a41) * k1[i])), u[i]);
            rhs(k4, tmp, p, fma(R_(c3), dt, t));
#pragma unroll
            for (int i = 0; i < N; i++)
                tmp[i] = fma(dt, fma(R_(a54), k4[i], fma(R_(a53), k3[i], fma(R_(a52), k2[i], R_(a51) * k1[i]))),
                             u[i]);
            rhs(k5, tmp, p, fma(R_(c4), dt, t));
#pragma unroll
            for (int i = 0; i < N; i++)
                tmp[i] = fma(dt,
                             fma(R_(a65), k5[i],
                                 fma(R_(a64), k4[i], fma(R_(a63), k3[i], fma(R_(a62), k2[i], R_(a61) * k1[i])))),
                             u[i]);
            rhs(k6, tmp, p, t + dt);
#pragma unroll
            for (int i = 0; i < N; i++)
                un[i] = fma(dt,
                            fma(R_(a76), k6[i],
                                fma(R_(a75), k5[i],
                                    fma(R_(a74), k4[i], fma(R_(a73), k3[i], fma(R_(a72), k2[i], R_(a71) * k1[i]))))),
                            u[i]);
            rhs(k7, un, p, t + dt);
            tnf += 6u;

            // ---- embedded error estimate: calculate_residuals + RMS norm ----
            real ss = R_(0.0);
#pragma unroll
            for (int i = 0; i < N; i++) {
                const real ut =
                    dt * fma(R_(bt7), k7[i],
                             fma(R_(bt6), k6[i],
                                 fma(R_(bt5), k5[i],
                                     fma(R_(bt4), k4[i], fma(R_(bt3), k3[i], fma(R_(bt2), k2[i], R_(bt1) * k1[i]))))));
                const real sc = fma(fmax(fabs(u[i]), fabs(un[i])), reltol, abstol);
                const real r = ut / sc;
                ss = fma(r, r, ss);
            }
            const real e2 = ss * (R_(1.0) / R_(N));  // EEst^2
            if (!(e2 == e2)) {  // NaN estimate: the reference rejects, dt turns NaN, solve aborts
                trej++;
                failed = 1;
                state = LANE_DONE;
                continue;
            }
            // ---- PI controller in log space: one log + one exp instead of two pow + sqrt ----
            // q = EEst^b1 / qold^b2 / gamma clamped to [1/qmax, 1/qmin]; we form 1/q directly.
            const real le = (e2 > R_(0.0)) ? R_(0.5) * log(e2) : neg_inf();  // log(EEst)
            if (e2 <= one_plus_eps) {  // sqrt(e2) <= 1  <=>  e2 <= nextafter(1)
                real qinv = R_(QMAX);
                if (e2 > R_(0.0)) {
                    qinv = R_(GAMMA) * exp(fma(R_(BETA2), lqold, -R_(BETA1) * le));
                    qinv = fmin(R_(QMAX), fmax(R_(QMIN), qinv));
                }
                lqold = fmax(le, R_(LOG_QOLDINIT));
                real tn = t + dt;
                if (fabs(tn - t1) < snap_tol) tn = t1;  // fixed_t_for_floatingpoint_error!
                t = tn;
#pragma unroll
                for (int i = 0; i < N; i++) { u[i] = un[i]; k1[i] = k7[i]; }
                dt = fmin(dtmax, dt * qinv);
                tacc++;
                if (!(t < t1)) state = LANE_DONE;
            } else {
                // step_reject_controller!: dt /= min(1/qmin, q11/gamma)
                const real f = fmax(R_(QMIN), R_(GAMMA) * exp(-R_(BETA1) * le));
                dt = dt * f;
                trej++;
            }
        }
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
    A.zero();
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
        A.nf += (unsigned long long)nsteps;
        A.nacc += (unsigned long long)nsteps;
        A.nfail += (unsigned)failed;
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
