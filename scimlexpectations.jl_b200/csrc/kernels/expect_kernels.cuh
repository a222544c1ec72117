// expect_kernels.cuh -- second part of the NVRTC translation unit, placed AFTER the user's model
// source (rhs / hmap / observable / diffusion are visible here and inline into the loops) and after
// expect_math.cuh.
//
// Replaces, for a whole batch per launch, what the reference does per trajectory on a CPU thread:
//   prob_func   src/expectation.jl:175-178 (MC :277-280, process noise :215-225 / :304-314)
//   solve       src/expectation.jl:191, :238, :290, :323  (Tsit5 / Euler-Maruyama)
//   output_func src/expectation.jl:180, :227 (g * pdf)  and :282, :316 (g)
//   reduction   src/expectation.jl:194 (dx .= ...)  and :293, :324 (mean(sol.u) -> sums here)
//
// Execution model: ONE trajectory per thread (two or three side by side in the Euler-Maruyama kernel), state
// and step control in registers; every block is resident (grid = SMs x resident blocks).  Adaptive step
// counts differ between lanes; instead of a nested per-trajectory loop (warp runs as long as its slowest
// lane) the loop body is "one RK step of my current trajectory", and idle lanes are re-loaded together once
// at least `refill_threshold` lanes of the warp are idle -- the reload (h, pdf, initial-dt heuristic, two
// extra RHS evaluations) costs about one step and would otherwise run under a divergent mask for every
// single finishing lane.  Which point a lane gets next is either fixed by the launch shape (static: lane l
// owns points l, l+T, l+2T, ...; sums reproducible) or decided at run time off a device counter (dynamic:
// no warp waits for work another one still owns; optionally with one row of sums per super-chunk of
// consecutive points so that the sums stay reproducible).  A point's own result -- its outputs, its step
// counters -- never depends on that choice.  Five instantiations per model: b200e_nodes, b200e_reduce
// (static), b200e_nodes_dyn, b200e_reduce_dyn (dynamic), b200e_reduce_gen (points drawn in the kernel).

#define B200E_FULL 0xffffffffu

#ifndef B200E_LOGPOLY_PDF
#define B200E_LOGPOLY_PDF 0
#endif

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
    double s = 0.0, tabulated = 1.0;
#pragma unroll
    for (int j = 0; j < NX; j++) {
        const int kind = P.pdf[j].kind;
        if (kind != 0) {
            const double xj = xs[j];
            const bool inside = (xj >= P.pdf[j].lo) && (xj <= P.pdf[j].hi);
            if (B200E_TABLE_PDF && kind == 4) {  // compiled in only when the model has a tabulated factor
                // tabulated factor (a KDE): pdf(kde, x) of KernelDensity = the quadratic B-spline through the grid
                // values (Interpolations' BSpline(Quadratic(Line(OnGrid()))), 0 outside).  The host prefiltered the
                // values into ntab + 2 coefficients (knot i <-> tab[i + 1]); three taps around the nearest knot:
                //   d = g - i in [-1/2, 1/2]:  (d - 1/2)^2 / 2 * c[i-1] + (3/4 - d^2) * c[i] + (d + 1/2)^2 / 2 * c[i+1]
                double v = 0.0;
                if (inside) {
                    const double g = (xj - P.pdf[j].lo) * P.pdf[j].inv_dx;
                    int i = (int)(g + 0.5);
                    i = i < P.pdf[j].ntab - 1 ? i : P.pdf[j].ntab - 1;
                    const double d = g - (double)i, dm = d - 0.5, dp = d + 0.5;
                    const double *c = P.pdf[j].tab + i;
                    const double cm = __ldg(c), c0 = __ldg(c + 1), cp = __ldg(c + 2);
                    v = fma(0.5 * dm * dm, cm, fma(0.75 - d * d, c0, 0.5 * dp * dp * cp));
                }
                tabulated *= v;
            } else if (B200E_LOGPOLY_PDF && kind == 5) {  // compiled in only when the model has such a factor
                // logpdf = c0 + c1 x + c2 x^2 + c3 log x + c4 log(1 - x) + c5 log(x)^2  (Exponential, Gamma, Beta,
                // LogNormal, Rayleigh, ... and their truncations; the coefficients sit in the sampler's fields)
                const double c3 = P.pdf[j].smp_a, c4 = P.pdf[j].smp_b, c5 = P.pdf[j].smp_c;
                double lp = fma(fma(P.pdf[j].inv_sigma, xj, P.pdf[j].mu), xj, P.pdf[j].logc);
                if (c3 != 0.0 || c5 != 0.0) {
                    const double lx = log(xj);
                    lp = fma(c5 != 0.0 ? fma(c5, lx, c3) : c3, lx, lp);   // no 0 * inf at x = 0
                }
                if (c4 != 0.0) lp = fma(c4, log1p(-xj), lp);
                s += inside ? lp : neg_inf();
            } else {
                const double z = (xj - P.pdf[j].mu) * P.pdf[j].inv_sigma;
                const double lp = fma(-0.5 * z, z, P.pdf[j].logc);
                s += inside ? lp : neg_inf();
            }
        }
    }
    return exp(s) * tabulated;
}

// GEN: the point is not loaded but drawn -- rand(d) of the MonteCarlo prob_func (expectation.jl:278) --
// from the counter-based stream, in registers (a separate instantiation: the sampler's code and
// registers stay out of the kernels that read x)
template <bool GEN>
__device__ __forceinline__ void load_point(const b200e_kparams &P, long long i, double *xs) {
    if (GEN) {
        b200e_smp::sample_point_fixed<NX>(P.pdf, P.gen_seed, (unsigned long long)(P.gen_first + i), xs);
    } else if (P.layout == 0) {
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
constexpr bool ACC_SMEM = (NACC * 8 + 4 * 8) * B200E_BLOCK <= 40960;  // the host sizes the block so that this holds
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

#ifndef B200E_NSAVE
#define B200E_NSAVE 0
#endif
constexpr int NSAVE = B200E_NSAVE;
constexpr int NPS = NSAVE > 0 ? NOUT / NSAVE : NOUT;  // outputs per save time

// weight + (store | accumulate) of outputs [first, first + count) of point i: output_func of the reference
template <bool REDUCE>
__device__ __forceinline__ void emit_values(const b200e_kparams &P, long long i, int first, int count, const real *out,
                                            double wgt, Accum &A) {
    for (int j = 0; j < count; j++) {
        double v = (double)out[j];
        if (P.weight_by_pdf) v *= wgt;
        const int k = first + j;
        if (REDUCE) {
            A.val(k) += v;
            A.val(NOUT + k) = fma(v, v, A.val(NOUT + k));
        } else {
            if (P.layout == 0) P.dx[i * NOUT + k] = v;
            else P.dx[(long long)k * P.ld + i] = v;
        }
    }
}

#if B200E_NSAVE == 0
// end-state observable g(sol[:, end], p)
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
#else
// time-sampled observable: slot k of g(sol, p) from the state at save time k (sol(t_k), sol[i, :] with saveat)
template <bool REDUCE>
__device__ __forceinline__ void emit_slot(const b200e_kparams &P, long long i, int k, real ts, const real *us,
                                          const real *p, double wgt, Accum &A) {
    real out[NPS];
    observable_at(k, ts, us, p, out);
    emit_values<REDUCE>(P, i, k * NPS, NPS, out, wgt, A);
}
// a trajectory that aborted never reached its remaining save times: their slots are NaN
template <bool REDUCE>
__device__ __forceinline__ void emit_missing(const b200e_kparams &P, long long i, int kfrom, double wgt, Accum &A) {
    real out[NPS];
#pragma unroll
    for (int j = 0; j < NPS; j++) out[j] = (real)__longlong_as_double(0x7ff8000000000000LL);
    for (int k = kfrom; k < NSAVE; k++) emit_values<REDUCE>(P, i, k * NPS, NPS, out, 1.0, A);
}
#endif

// warp shuffle + shared-memory block reduction; only per-block partials reach HBM.  Counters are
// folded in both modes; the value accumulators only in reduce mode.
template <bool REDUCE>
__device__ __forceinline__ void block_reduce_store(const b200e_kparams &P, Accum &A) {
    __shared__ double sacc[NWARPS][NACC];
    __shared__ unsigned long long scnt[NWARPS][4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long cnt[4];
#pragma unroll
    for (int k = 0; k < 4; k++) cnt[k] = A.cnt(k);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
        for (int k = 0; k < 4; k++) cnt[k] += __shfl_down_sync(B200E_FULL, cnt[k], off);
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 4; k++) scnt[warp][k] = cnt[k];
    }
    if (REDUCE) {
        // one column at a time: a wide observable (2 nout columns) must not sit in registers all at once
#pragma unroll 1
        for (int k = 0; k < NACC; k++) {
            double v = A.val(k);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(B200E_FULL, v, off);
            if (lane == 0) sacc[warp][k] = v;
        }
    }
    __syncthreads();
    if (REDUCE) {
        for (int k = threadIdx.x; k < NACC; k += B200E_BLOCK) {  // NACC may exceed a small block
            double t = 0.0;
            for (int w = 0; w < NWARPS; w++) t += sacc[w][k];
            P.partials[(long long)blockIdx.x * NACC + k] = t;
        }
    }
    if (threadIdx.x < 4) {
        const int k = threadIdx.x;
        unsigned long long t = 0ull;
        for (int w = 0; w < NWARPS; w++) t += scnt[w][k];
        P.cpartials[(long long)blockIdx.x * 4 + k] = t;
    }
    // ---- the last block to get here folds the partials of the whole launch (fixed order: the result
    // depends on the launch shape only) and re-arms the launch's counters, so a single-launch call needs
    // no finalize kernels and no memset ----
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(P.work_counter + 1, 1ull) == (unsigned long long)(gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x == 0) { P.work_counter[0] = 0ull; P.work_counter[1] = 0ull; }
    // one warp per column (lane-strided compensated sum, then the shuffle tree): no block barriers
    const long long rows = gridDim.x;
    const int nsum = (REDUCE && P.final_sums) ? NACC : 0;
    const int ncol = nsum + (P.final_counts ? 4 : 0);
    for (int c = warp; c < ncol; c += NWARPS) {
        if (c < nsum) {
            double s = 0.0, comp = 0.0;  // Neumaier
            for (long long r = lane; r < rows; r += 32) {
                const double v = __ldcg(P.partials + r * NACC + c);
                const double t = s + v;
                comp += (fabs(s) >= fabs(v)) ? ((s - t) + v) : ((v - t) + s);
                s = t;
            }
            s += comp;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(B200E_FULL, s, off);
            if (lane == 0) P.final_sums[c] = s;
        } else {
            const int k = c - nsum;
            unsigned long long s = 0ull;
            for (long long r = lane; r < rows; r += 32) s += __ldcg(P.cpartials + r * 4 + k);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(B200E_FULL, s, off);
            if (lane == 0) P.final_counts[k] = s;
        }
    }
}

#if B200E_SOLVER != 1
// ------------------------------------------------------------------------------------------
// The adaptive solvers share one driver (adaptive_run: refill phase + step phase, work distribution,
// accumulators) and differ in how a lane starts a trajectory and takes one attempted step:
//   B200E_SOLVER 0  Tsit5 (Tsitouras 2011) with OrdinaryDiffEq's defaults: Hairer-Wanner initial step,
//                   PI controller beta1=7/50 beta2=2/25 gamma=9/10 qmin=1/5 qmax=10 qoldinit=1e-4, RMS norm;
//   B200E_SOLVER 2  LambaEM (adaptive Euler-Maruyama, the solver of reference test/processnoise.jl:15)
//                   driven by the KKL path, StochasticDiffEq's defaults (see lamba_step).
// Constants (KC / K_), the table-driven log / exp and the reciprocal live in expect_math.cuh.
// ------------------------------------------------------------------------------------------
__shared__ b200e_tabs sh_tab;

// Where a lane is: EMPTY (no trajectory), ACTIVE (stepping), DONE (reached t1), FAILED (aborted in a loop header: a
// finished trajectory in its own state rather than DONE plus a flag -- the flag was one more value live across the step
// loop and cost a save / restore pair of moves per attempt).  (Tried: the state in the top bits of the accepted-step
// counter, no separate variable at all -- ptxas spends the saved instructions on moves around the counter.)
enum { LANE_EMPTY = 0, LANE_ACTIVE = 1, LANE_DONE = 2, LANE_FAILED = 3 };
template <class LaneT> __device__ __forceinline__ bool lane_active(const LaneT &L) { return L.state == LANE_ACTIVE; }
template <class LaneT> __device__ __forceinline__ bool lane_empty(const LaneT &L) { return L.state == LANE_EMPTY; }
template <class LaneT> __device__ __forceinline__ bool lane_finished(const LaneT &L) { return L.state >= LANE_DONE; }
template <class LaneT> __device__ __forceinline__ unsigned lane_failed(const LaneT &L) { return L.state == LANE_FAILED ? 1u : 0u; }
template <class LaneT> __device__ __forceinline__ unsigned lane_naccept(const LaneT &L) { return L.nacc; }
template <class LaneT> __device__ __forceinline__ void lane_set_empty(LaneT &L) { L.nacc = 0; L.state = LANE_EMPTY; }
template <class LaneT> __device__ __forceinline__ void lane_abort(LaneT &L) { L.state = LANE_FAILED; }
template <class LaneT> __device__ __forceinline__ void lane_start(LaneT &L, const bool active) { L.nacc = 0; L.state = active ? LANE_ACTIVE : LANE_DONE; }
template <class LaneT> __device__ __forceinline__ void lane_accept(LaneT &L, const bool done) { L.nacc++; if (done) L.state = LANE_DONE; }

#if B200E_SOLVER == 0
constexpr unsigned NF_PER_ATTEMPT = 6u;  // RHS evaluations per attempted step

// prob_func + integrator init for one lane: x -> (u0, p) via h, pdf weight, initial dt, FSAL k1
template <bool GEN>
__device__ __forceinline__ void start_trajectory(const b200e_kparams &P, long long i, real *u, real *p,
                                                 real *k1, real &t, real &dt, real &le2old, double &wgt,
                                                 unsigned &nf) {
    double xs[NX];
    load_point<GEN>(P, i, xs);
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
    // d1 = Inf over a finite d0 (an infinite parameter or state derivative): the reference's d0 / d1 is exactly 0
    // and the solve takes ode_determine_initdt's "singular" exit below after two evaluations; the reciprocal
    // above would turn Inf into NaN, which fmin drops (the regular path, one evaluation more)
    if (s1 == (real)__longlong_as_double(0x7ff0000000000000LL) && s0 < s1) dt0 = R_(0.0);
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
        else dt1 = exp_small(sh_tab, fma(R_(-0.1), log_pos(sh_tab, m2), -K_(ln10_04)));  // 10^(-(2 + log10 dmax)/5)
        dt = fmin(fmin(R_(100.0) * dt0, dt1), dtmax);
        nf += 3u;
    }
    // Tsit5 initialize!: fsalfirst = f(uprev, p, t) (same value as f0; the reference evaluates it
    // again, so it is counted again)
#pragma unroll
    for (int j = 0; j < N; j++) k1[j] = f0[j];
    le2old = K_(le2_qoldinit_n);
}

// One attempted Tsit5 step of the lane's trajectory (perform_step! + stepsize_controller! +
// step_accept/reject_controller! + the loopheader! checks of the reference's integrator loop).
struct Lane {
    real u[N], k1[N], p[B200E_DIM(NP)];
    real t, dt, le2old;
    int state;
    unsigned nf0, nacc, nrej;  // nf = nf0 (initial-step evaluations) + 6 * (nacc + nrej)
    int ksave;                 // time-sampled observables: next save slot
};

// what the accept branch needs to emit save slots (unused, and compiled away, when B200E_NSAVE == 0)
struct SaveCtx {
    const b200e_kparams &P;
    long long point;
    double wgt;
    Accum &A;
};

template <bool REDUCE>
__device__ __forceinline__ void tsit5_step(Lane &L, const real t1, const real dtmax,
                                           const real abstol, const real reltol, const real snap_tol,
                                           const int maxiters, const real one_plus_eps, const SaveCtx &S) {
    real *u = L.u, *k1 = L.k1, *p = L.p;
    const real t = L.t;
    // ---- loopheader!: a NaN error estimate in the previous attempt, the iteration cap, or a step
    // below the resolution of t abort the trajectory (dt <= dtmax holds where dt is proposed) ----
    const real dt = pmin(L.dt, t1 - t);  // clip to the end point
    const real tnext = t + dt;
    if ((int)(L.nacc + L.nrej) >= maxiters || same_value(tnext, t)) {  // (a NaN estimate left dt = 0: tnext == t)
        lane_abort(L);
        return;
    }
    // ---- perform_step!: 6 new stages, FSAL.  The stages are kept pre-multiplied by dt (kd_j = dt k_j):
    // every stage sum is then a chain of DFMAs with ONE register factor and one constant-bank factor.
    // A DFMA with three distinct register sources issues every 3 cycles on sm_100, one with two every
    // 2 cycles (profiles/r1_dp_operands_ubench.txt), and "u + dt * sum" would be of the slow kind. ----
    real kd1[N], kd2[N], kd3[N], kd4[N], kd5[N], kd6[N], k[N], k7[N], tmp[N], un[N];
#pragma unroll
    for (int i = 0; i < N; i++) { kd1[i] = dt * k1[i]; tmp[i] = fma(K_(a21), kd1[i], u[i]); }
    rhs(k, tmp, p, fma(K_(c1), dt, t));
#pragma unroll
    for (int i = 0; i < N; i++) { kd2[i] = dt * k[i]; tmp[i] = fma(K_(a32), kd2[i], fma(K_(a31), kd1[i], u[i])); }
    rhs(k, tmp, p, fma(K_(c2), dt, t));
#pragma unroll
    for (int i = 0; i < N; i++) {
        kd3[i] = dt * k[i];
        tmp[i] = fma(K_(a43), kd3[i], fma(K_(a42), kd2[i], fma(K_(a41), kd1[i], u[i])));
    }
    rhs(k, tmp, p, fma(K_(c3), dt, t));
#pragma unroll
    for (int i = 0; i < N; i++) {
        kd4[i] = dt * k[i];
        tmp[i] = fma(K_(a54), kd4[i], fma(K_(a53), kd3[i], fma(K_(a52), kd2[i], fma(K_(a51), kd1[i], u[i]))));
    }
    rhs(k, tmp, p, fma(K_(c4), dt, t));
#pragma unroll
    for (int i = 0; i < N; i++) {
        kd5[i] = dt * k[i];
        tmp[i] = fma(K_(a65), kd5[i],
                     fma(K_(a64), kd4[i], fma(K_(a63), kd3[i], fma(K_(a62), kd2[i], fma(K_(a61), kd1[i], u[i])))));
    }
    rhs(k, tmp, p, tnext);
#pragma unroll
    for (int i = 0; i < N; i++) {
        kd6[i] = dt * k[i];
        un[i] = fma(K_(a76), kd6[i],
                    fma(K_(a75), kd5[i],
                        fma(K_(a74), kd4[i], fma(K_(a73), kd3[i], fma(K_(a72), kd2[i], fma(K_(a71), kd1[i], u[i]))))));
    }
    rhs(k7, un, p, tnext);

    // ---- embedded error estimate: calculate_residuals + RMS norm, kept squared ----
    real ss = R_(0.0);
#pragma unroll
    for (int i = 0; i < N; i++) {
        const real ut = fma(K_(bt7), dt * k7[i],
                            fma(K_(bt6), kd6[i],
                                fma(K_(bt5), kd5[i],
                                    fma(K_(bt4), kd4[i], fma(K_(bt3), kd3[i], fma(K_(bt2), kd2[i], K_(bt1) * kd1[i]))))));
        const real sc = fma(pmax(abs_bits(u[i]), abs_bits(un[i])), reltol, abstol);
        const real r = ut * rcp_pos(sc);
        ss = fma(r, r, ss);
    }
    // 1/N is a DFMA immediate when N is a power of two
    // EEst^2 -- times N, undivided, when N is a power of two (lit::e2_scaled in expect_math.cuh)
    const real e2 = lit::e2_scaled ? ss : ss * K_(inv_n);
    const bool accept = le_nonneg(e2, one_plus_eps * R_(lit::e2_scale));  // sqrt(e2) <= 1  <=>  e2 <= nextafter(1); false for a NaN
#if B200E_NSAVE > 0
    // before the controller, while the stages are still live anyway (keeps them out of its register budget)
    if (accept) {
        const real left0 = t1 - tnext;
        const real tnew = lt_nonneg(abs_bits(left0), snap_tol) ? t1 : tnext;  // the same snap as below
        // saveat / sol(t): every save time inside (t, t_new] from Tsit5's free 4th-order interpolant
        //   u(t + th dt) = u + sum_i b_i(th) kd_i,  b_1 = th (r11 + th (r12 + th (r13 + th r14))),
        //   b_i = th^2 (ri2 + th (ri3 + th ri4))   (Tsitouras 2011; OrdinaryDiffEq's tsit_interpolants);
        // a save time at the end of the step takes the step's own end state
        while (L.ksave < NSAVE && (real)S.P.save_t[L.ksave] <= tnew) {
            const real ts = (real)S.P.save_t[L.ksave];
            real us[N];
            if (ts >= tnew) {
#pragma unroll
                for (int i = 0; i < N; i++) us[i] = un[i];
            } else {
                const real th = (ts - t) / dt, th2 = th * th;
                const real b1 = th * (R_(1.0) + th * (R_(-2.763706197274826) + th * (R_(2.9132554618219126) + th * R_(-1.0530884977290216))));
                const real b2 = th2 * (R_(0.13169999999999998) + th * (R_(-0.2234) + th * R_(0.1017)));
                const real b3 = th2 * (R_(3.9302962368947516) + th * (R_(-5.941033872131505) + th * R_(2.490627285651253)));
                const real b4 = th2 * (R_(-12.411077166933676) + th * (R_(30.33818863028232) + th * R_(-16.548102889244902)));
                const real b5 = th2 * (R_(37.50931341651104) + th * (R_(-88.1789048947664) + th * R_(47.37952196281928)));
                const real b6 = th2 * (R_(-27.896526289197286) + th * (R_(65.09189467479366) + th * R_(-34.87065786149661)));
                const real b7 = th2 * (R_(1.5) + th * (R_(-4.0) + th * R_(2.5)));
#pragma unroll
                for (int i = 0; i < N; i++)
                    us[i] = u[i] + (b1 * kd1[i] + b2 * kd2[i] + b3 * kd3[i] + b4 * kd4[i] + b5 * kd5[i] + b6 * kd6[i] +
                                    b7 * (dt * k7[i]));
            }
            emit_slot<REDUCE>(S.P, S.point, L.ksave, ts, us, p, S.wgt, S.A);
            L.ksave++;
        }
    }
#endif

    // ---- PI controller in log space on e2: one log + one exp instead of two pow + sqrt ----
    // accept:  1/q = gamma * qold^b2 / EEst^b1  clamped to [qmin, qmax]
    // reject:  dt *= max(qmin, gamma / EEst^b1)  (< 1, so the same clamp applies)
    // No special cases: log_pos maps e2 = 0 / subnormal to about -711 and e2 = +inf / NaN to +708, for
    // which the clamps on 1/q decide alone exactly as the reference's (EEst == 0 -> q = 1/qmax;
    // EEst = Inf -> dt/5), and |b2*le2old - b1*le2| <= 51 stays far inside exp_small's range.
    const real le2 = log_pos(sh_tab, e2);
    const real hist = accept ? L.le2old : R_(lit::ln_n);
    real qinv = exp_small(sh_tab, fma(K_(hb2), hist, fma(-K_(hb1), le2, K_(ln_gamma_n))));  // gamma * e2old^hb2 / e2^hb1
    qinv = pmin(R_(lit::qmax), pmax(K_(qmin), qinv));
    L.dt = dt * qinv;
#if B200E_CLIP_DTMAX  // compiled in only when the model sets dtmax < t1 - t0
    L.dt = pmin(dtmax, L.dt);
#endif
    if (accept) {
        L.le2old = max_with_negative(le2, R_(lit::le2_qoldinit_n));  // qold = max(EEst, 1e-4); le2 is never NaN
        // fixed_t_for_floatingpoint_error! (|t_new - t1| < 100 ulp snaps to t1) and the loop condition !(t < t1) in
        // one test: t1 - t_new < snap_tol, negative values included (the time of a finished trajectory is not used)
        const real left = t1 - tnext;
        const bool done = lt_any_pos(left, snap_tol);
        L.t = done ? t1 : tnext;
#pragma unroll
        for (int i = 0; i < N; i++) { u[i] = un[i]; k1[i] = k7[i]; }
        lane_accept(L, done);
    } else {
        // NaN estimate: the reference counts a rejection, dt turns NaN and the solve aborts at the next
        // loop header -- same here through dt = 0 (t + dt == t aborts there; tested on this rarely taken side only)
        if (!(e2 == e2)) L.dt = R_(0.0);
        L.nrej++;
    }
}

#else  // B200E_SOLVER == 2
// ------------------------------------------------------------------------------------------
// LambaEM: Euler-Maruyama with the Lamba / Rackauckas error estimate and the SDE PI controller
// (StochasticDiffEq; not vendored with the reference, restated from package knowledge -- the oracle's
// lamba_solve carries the same statement and the "parity unpinned" note):
//   K = u + dt f(u,t);  L = g(K, t+dt);  u' = K + L dW,  dW = W(t+dt) - W(t)   (NoiseFunction, split step)
//   Ed = dt (f(K,t+dt) - f(u,t)) / 2;   En = (g(u + L sqrt(dt), t) - L) / sqrt(dt) * dW^2 / 2
//   EEst = RMS_i( (|Ed_i| + |En_i|) / (abstol + max(|u_i|,|u'_i|) reltol) )
//   beta1 = 7/10, beta2 = 2/5, gamma = 9/10, qmin = 1/5, qmax = 1.125, qoldinit = 1e-4.
// W(t) = sum_k Z_k sqrt(2) sin((k-1/2) pi t) / ((k-1/2) pi) is needed at arbitrary t here (no table as in
// the fixed-step kernel): ONE sine (sin_halfpi, 15 FP64 instructions) and the three-term recurrence
//   s_{k+1} = 2 cos(pi t) s_k - s_{k-1},  s_k = sin((k-1/2) pi t),  s_0 = -s_1
// give all NK modes in 2 DFMAs each; the reference evaluates NK sines per query (src/expectation.jl:217-220).
// ------------------------------------------------------------------------------------------
constexpr unsigned NF_PER_ATTEMPT = 2u;  // drift evaluations per attempted step
constexpr int NK = B200E_NKKL;

struct Lane {
    real u[N], zc[NK], p[B200E_DIM(NP)];  // zc_k = Z_k sqrt(2) / ((k-1/2) pi)
    real t, dt, le2old, wcur;
    int state;
    unsigned nf0, nacc, nrej;
    int ksave;
};
struct SaveCtx {
    const b200e_kparams &P;
    long long point;
    double wgt;
    Accum &A;
};

__device__ __forceinline__ real kkl_w(const real *zc, const real t) {
    const real s = sin_halfpi(t);
    const real c2 = fma(R_(-4.0) * s, s, R_(2.0));  // 2 cos(pi t) = 2 - 4 sin^2(pi t / 2)
    real sp = -s, sk = s, w = R_(0.0);
#pragma unroll
    for (int k = 0; k < NK; k++) {
        w = fma(zc[k], sk, w);
        const real sn = fma(c2, sk, -sp);
        sp = sk;
        sk = sn;
    }
    return w;
}

// prob_func (src/expectation.jl:215-225: Z, p = h(x, u0, p); u0 stays nominal) + sde_determine_initdt
template <bool GEN>
__device__ __forceinline__ void start_trajectory(const b200e_kparams &P, long long i, Lane &L, double &wgt) {
    double xs[NX];
    load_point<GEN>(P, i, xs);
    wgt = P.weight_by_pdf ? pdf_weight(P, xs) : 1.0;
    real xr[NX], u0n[N], pn[B200E_DIM(NP)], z[NK];
#pragma unroll
    for (int j = 0; j < NX; j++) xr[j] = (real)xs[j];
#pragma unroll
    for (int j = 0; j < N; j++) { u0n[j] = (real)P.u0_nom[j]; L.u[j] = u0n[j]; }
#pragma unroll
    for (int j = 0; j < NP; j++) { pn[j] = (real)P.p_nom[j]; L.p[j] = pn[j]; }
#pragma unroll
    for (int k = 0; k < NK; k++) z[k] = R_(0.0);
    hmap(xr, u0n, pn, z, L.p);
#pragma unroll
    for (int k = 0; k < NK; k++) L.zc[k] = z[k] * R_(1.4142135623730951 / ((k + 0.5) * 3.14159265358979323846));
    const real *u = L.u, *p = L.p;
    const real t = (real)P.t0;
    const real abstol = (real)P.abstol, reltol = (real)P.reltol, dtmax = (real)P.dtmax;
    // sde_determine_initdt on squared norms; max(|f + 3g|, |f - 3g|) = |f| + 3|g| exactly
    real f0[N], g0[N], isk[N];
    rhs(f0, u, p, t);
    diffusion(g0, u, p, t);
    real s0 = R_(0.0), s1 = R_(0.0);
#pragma unroll
    for (int j = 0; j < N; j++) {
        isk[j] = rcp_pos(fma(fabs(u[j]), reltol, abstol));
        g0[j] = R_(3.0) * fabs(g0[j]);
        const real a = u[j] * isk[j], b = (fabs(f0[j]) + g0[j]) * isk[j];
        s0 = fma(a, a, s0);
        s1 = fma(b, b, s1);
    }
    const real tiny2 = R_(1e-10) * R_(N);
    real dt0 = (s0 < tiny2 || s1 < tiny2) ? R_(1e-6) : R_(0.01) * sqrt_pos(s0 * rcp_pos(s1));
    dt0 = fmin(dt0, dtmax);
    real u1[N], f1[N], g1[N];
#pragma unroll
    for (int j = 0; j < N; j++) u1[j] = fma(dt0, f0[j], u[j]);
    rhs(f1, u1, p, t + dt0);
    diffusion(g1, u1, p, t + dt0);
    real s2 = R_(0.0);
#pragma unroll
    for (int j = 0; j < N; j++) {
        // max(|g0 - g1|, |g0 + g1|) = |g0| + |g1|;  max(|df + dg|, |df - dg|) = |df| + dg
        const real d = (fabs(f1[j] - f0[j]) + (g0[j] + R_(3.0) * fabs(g1[j]))) * isk[j];
        s2 = fma(d, d, s2);
    }
    const real idt0 = rcp_pos(dt0);
    const real m2 = fmax(s1, s2 * idt0 * idt0) * R_(lit::inv_n);  // max(d1, d2)^2
    // order 1/2: dt1 = 10^(-(2 + log10 dmax) / (order + 1/2)) = 0.01 / dmax
    const real dt1 = (m2 <= R_(1e-30)) ? fmax(R_(1e-6), dt0 * R_(1e-3)) : R_(0.01) * rcp_pos(sqrt_pos(m2));
    L.t = t;
    L.dt = fmin(fmin(R_(100.0) * dt0, dt1), dtmax);
    L.wcur = kkl_w(L.zc, t);
    L.le2old = K_(le2_qoldinit_n);
    L.nf0 += 2u;
}

template <bool REDUCE>
__device__ __forceinline__ void lamba_step(Lane &L, const real t1, const real dtmax, const real abstol,
                                           const real reltol, const real snap_tol, const int maxiters,
                                           const real one_plus_eps) {
    real *u = L.u, *p = L.p;
    const real t = L.t;
    const real dt = pmin(L.dt, t1 - t);  // modify_dt_for_tstops!
    const real tnext = t + dt;
    if ((int)(L.nacc + L.nrej) >= maxiters || same_value(tnext, t)) {  // (a NaN estimate left dt = 0: tnext == t)
        lane_abort(L);
        return;
    }
    real f0[N], K[N], Lg[N], un[N], f1[N], v[N], g1[N];
    rhs(f0, u, p, t);
#pragma unroll
    for (int i = 0; i < N; i++) K[i] = fma(dt, f0[i], u[i]);
    diffusion(Lg, K, p, tnext);
    const real wn = kkl_w(L.zc, tnext), dW = wn - L.wcur;
#pragma unroll
    for (int i = 0; i < N; i++) un[i] = fma(Lg[i], dW, K[i]);
    rhs(f1, K, p, tnext);
    real sq, isq;
    sqrt_rsqrt_pos(dt, sq, isq);
#pragma unroll
    for (int i = 0; i < N; i++) v[i] = fma(Lg[i], sq, u[i]);
    diffusion(g1, v, p, t);
    const real hdt = R_(0.5) * dt, hw = R_(0.5) * dW * dW * isq;  // both >= 0
    real ss = R_(0.0);
#pragma unroll
    for (int i = 0; i < N; i++) {
        // |Ed| + |En| = |f1 - f0| dt/2 + |g1 - L| dW^2 / (2 sqrt(dt))
        const real num = fma(abs_bits(g1[i] - Lg[i]), hw, abs_bits(f1[i] - f0[i]) * hdt);
        const real sc = fma(pmax(abs_bits(u[i]), abs_bits(un[i])), reltol, abstol);
        const real r = num * rcp_pos(sc);
        ss = fma(r, r, ss);
    }
    const real e2 = lit::e2_scaled ? ss : ss * R_(lit::inv_n);                 // EEst^2 (times N: see tsit5_step)
    const bool accept = le_nonneg(e2, one_plus_eps * R_(lit::e2_scale));      // false for a NaN
    // PI controller in log space on e2 (see tsit5_step): EEst^b1 = e2^(b1/2)
    const real le2 = log_pos(sh_tab, e2);
    const real hist = accept ? L.le2old : R_(lit::ln_n);
    real qinv = exp_small(sh_tab, fma(K_(sde_hb2), hist, fma(-K_(sde_hb1), le2, K_(sde_ln_gamma_n))));  // gamma qold^b2 / EEst^b1
    qinv = pmin(R_(1.125), pmax(K_(qmin), qinv));
    L.dt = dt * qinv;
#if B200E_CLIP_DTMAX
    L.dt = pmin(dtmax, L.dt);
#endif
    if (accept) {
        L.le2old = max_with_negative(le2, R_(lit::le2_qoldinit_n));
        const real left = t1 - tnext;
        const bool done = lt_any_pos(left, snap_tol);  // fixed_t_for_floatingpoint_error! and !(t < t1), see tsit5_step
        L.t = done ? t1 : tnext;
#pragma unroll
        for (int i = 0; i < N; i++) u[i] = un[i];
        L.wcur = wn;
        lane_accept(L, done);
    } else {
        if (!(e2 == e2)) L.dt = R_(0.0);   // NaN estimate: abort at the next loop header
        L.nrej++;
    }
}
#endif  // solver-specific start / step

template <bool REDUCE, bool DYN, bool GEN>
__device__ __forceinline__ void tsit5_run(const b200e_kparams &P) {
    // Work distribution.  Static: lane l owns points l, l+T, l+2T, ... (T = grid * block).  Dynamic:
    // whenever a warp refills, its idle lanes take the next consecutive points off the launch's counter;
    // which warp solves which point is decided at run time, so a warp the SM favoured simply takes more (resident warps do not advance at equal pace, and with fixed
    // ownership the last ones finish alone, latency-bound).  A launch with at most one point per lane
    // has nothing to balance: the host launches the static instantiation for those (no atomics in the
    // small Koopman refinement rounds) and whenever the model asks for B200E_SCHED_STATIC.
    const long long T = (long long)gridDim.x * B200E_BLOCK;
    constexpr bool dyn = DYN;
    long long next = dyn ? P.npts : (long long)blockIdx.x * B200E_BLOCK + threadIdx.x;
    long long cur = -1;

    Lane L;
    L.t = R_(0.0); L.dt = R_(0.0); L.le2old = R_(0.0);
#if B200E_SOLVER == 2
    L.wcur = R_(0.0);
#endif
    L.nf0 = 0; L.nrej = 0; L.ksave = 0;
    lane_set_empty(L);
    double wgt = 1.0;
    Accum A;
    A.init();
#if !B200E_FP32
    for (int j = threadIdx.x; j < TAB_N; j += B200E_BLOCK) tab_fill(sh_tab, j);
    __syncthreads();
#endif

    const real t1 = (real)P.t1, dtmax = (real)P.dtmax;
    const real abstol = (real)P.abstol, reltol = (real)P.reltol;
    const real snap_tol = (real)P.snap_tol;
    const int maxiters = (int)P.maxiters;
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

    bool more = dyn;  // dynamic mode, warp-uniform: the launch still has unassigned points
    const unsigned lt_mask = (1u << (threadIdx.x & 31)) - 1u;

    // Reproducible dynamic schedule (B200E_SCHED_DYNAMIC_REPRODUCIBLE): the warp takes SUPER-CHUNKS of
    // P.super_chunks x 32 consecutive points off the counter, starts 32 trajectories at a time, sums every
    // finished chunk over its lanes with a fixed shuffle tree, adds the chunk sums in chunk order, and
    // writes ONE row of sums per super-chunk.  A row is a function of the points of its super-chunk only
    // and the rows are folded in index order afterwards: the result does not depend on which warp solved
    // what, nor on the launch shape.
    constexpr bool repro = B200E_REPRO && REDUCE && DYN;
    __shared__ double sh_wacc[repro ? NWARPS : 1][repro ? NACC : 1];
    long long sc_next = 0, sc_row = -1;
    int sc_left = 0;
    if (repro) {
        thr_eff = 32;
        tune = false;
        if ((threadIdx.x & 31) < NACC) sh_wacc[threadIdx.x >> 5][threadIdx.x & 31] = 0.0;
        for (int k = 32 + (threadIdx.x & 31); k < NACC; k += 32) sh_wacc[threadIdx.x >> 5][k] = 0.0;
        __syncwarp();
    }
    for (;;) {
        // ---- refill phase: runs only when the set of stepping lanes has changed ----
        const bool wants = lane_finished(L) ||
                           (lane_empty(L) && (dyn ? (more || (repro && sc_left > 0)) : next < P.npts));
        const unsigned idle_m = __ballot_sync(B200E_FULL, wants);
        unsigned act_m = __ballot_sync(B200E_FULL, lane_active(L));
        if ((idle_m | act_m) == 0u) break;
        const int live = __popc(idle_m | act_m);
        const int thr = thr_eff < live ? thr_eff : live;
        if (__popc(idle_m) >= thr) {  // warp-uniform
            const unsigned done_m = __ballot_sync(B200E_FULL, lane_finished(L));
            if (tune && done_m == B200E_FULL) {  // first full generation finished: measure its imbalance
                const unsigned st = lane_naccept(L) + L.nrej;
                const unsigned mx = __reduce_max_sync(B200E_FULL, st), sm = __reduce_add_sync(B200E_FULL, st);
                if (mx * 128u > sm * 5u) thr_eff = 8;  // max > 1.25 * mean
                tune = false;
            }
#if B200E_NSAVE > 0
            if (repro && done_m != 0u) {
                // time-sampled observables: the accept branch has put the (weighted) slot values of the lane's ONE
                // trajectory of this chunk, and their squares, into the lane's accumulator columns (unused otherwise
                // under this schedule); aborted trajectories add their NaN slots now.  Chunk sums in a fixed order,
                // then the columns are cleared for the next chunk.
                if (lane_finished(L)) emit_missing<REDUCE>(P, cur, L.ksave, wgt, A);
#pragma unroll 1
                for (int k = 0; k < NACC; k++) {
                    double v = A.val(k);
                    A.val(k) = 0.0;
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(B200E_FULL, v, off);
                    if ((threadIdx.x & 31) == 0) sh_wacc[threadIdx.x >> 5][k] += v;
                }
            }
#else
            if (repro && done_m != 0u) {
                // chunk sum in a fixed order: lanes without a point contribute zeros
                real out[NOUT];
#pragma unroll
                for (int k = 0; k < NOUT; k++) out[k] = R_(0.0);
                if (lane_finished(L)) observable(L.u, L.p, (real)P.t1, out);
#pragma unroll 1
                for (int k = 0; k < NOUT; k++) {
                    double v = lane_finished(L) ? (double)out[k] : 0.0;
                    if (P.weight_by_pdf) v *= wgt;
                    double v2 = v * v;
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
                        v += __shfl_xor_sync(B200E_FULL, v, off);
                        v2 += __shfl_xor_sync(B200E_FULL, v2, off);
                    }
                    if ((threadIdx.x & 31) == 0) {
                        sh_wacc[threadIdx.x >> 5][k] += v;
                        sh_wacc[threadIdx.x >> 5][NOUT + k] += v2;
                    }
                }
            }
#endif
            if (lane_finished(L)) {
#if B200E_NSAVE > 0
                if (!repro) emit_missing<REDUCE>(P, cur, L.ksave, wgt, A);
#else
                if (!repro) emit<REDUCE>(P, cur, L.u, L.p, wgt, A);
#endif
                const unsigned nacc = lane_naccept(L);
                if (!REDUCE && P.steps) P.steps[cur] = (int)(nacc + L.nrej);
                A.count(L.nf0 + NF_PER_ATTEMPT * (nacc + L.nrej), nacc, L.nrej, lane_failed(L));
                lane_set_empty(L);
            }
            if (repro) {
                if (sc_left == 0 && sc_row >= 0) {  // the super-chunk is complete: its row leaves the warp
                    __syncwarp();
                    for (int k = threadIdx.x & 31; k < NACC; k += 32) {
                        P.chunk_rows[sc_row * NACC + k] = sh_wacc[threadIdx.x >> 5][k];
                        sh_wacc[threadIdx.x >> 5][k] = 0.0;
                    }
                    __syncwarp();
                    sc_row = -1;
                }
                next = P.npts;
                if (sc_left == 0 && more) {
                    const long long span = (long long)P.super_chunks * 32;
                    unsigned long long base = 0ull;
                    if ((threadIdx.x & 31) == 0) base = atomicAdd(P.work_counter, (unsigned long long)span);
                    base = __shfl_sync(B200E_FULL, base, 0);
                    more = (long long)base + span < P.npts;
                    if ((long long)base < P.npts) {
                        const long long left = P.npts - (long long)base;
                        sc_row = (long long)base / span;
                        sc_next = (long long)base;
                        sc_left = (int)((left < span ? left : span) + 31) / 32;
                    }
                }
                if (sc_left > 0) {
                    next = sc_next + (threadIdx.x & 31);
                    sc_next += 32;
                    sc_left--;
                }
            } else if (dyn) {
                next = P.npts;
                if (more) {  // one atomic per refill: the warp takes the next popc(idle) points of the launch
                    unsigned long long base = 0ull;
                    if ((threadIdx.x & 31) == 0) base = atomicAdd(P.work_counter, (unsigned long long)__popc(idle_m));
                    base = __shfl_sync(B200E_FULL, base, 0);
                    if (wants) next = (long long)base + __popc(idle_m & lt_mask);
                    more = (long long)base + __popc(idle_m) < P.npts;
                }
            }
            if (wants && next < P.npts) {
                cur = next;
                next = dyn ? P.npts : next + T;
                L.nf0 = 0; L.nrej = 0;
#if B200E_SOLVER == 2
                start_trajectory<GEN>(P, cur, L, wgt);
#else
                start_trajectory<GEN>(P, cur, L.u, L.p, L.k1, L.t, L.dt, L.le2old, wgt, L.nf0);
#endif
                lane_start(L, L.t < t1);
#if B200E_NSAVE > 0
                // save times at t0 hold the initial state (save_start)
                for (L.ksave = 0; L.ksave < NSAVE && (real)P.save_t[L.ksave] <= L.t; L.ksave++)
                    emit_slot<REDUCE>(P, cur, L.ksave, (real)P.save_t[L.ksave], L.u, L.p, wgt, A);
#endif
            }
            act_m = __ballot_sync(B200E_FULL, lane_active(L));
        }
        if (act_m == 0u) continue;
        // ---- step phase: a tight loop, one vote per step, until some lane finishes (a lane only ever goes from
        // ACTIVE to DONE in here, so "nobody finished" is one all-vote on a per-lane predicate) ----
        const bool stepping = lane_active(L);
        do {
#if B200E_SOLVER == 2
            if (lane_active(L)) lamba_step<REDUCE>(L, t1, dtmax, abstol, reltol, snap_tol, maxiters, one_plus_eps);
#else
            if (lane_active(L)) tsit5_step<REDUCE>(L, t1, dtmax, abstol, reltol, snap_tol, maxiters, one_plus_eps, SaveCtx{P, cur, wgt, A});
#endif
        } while (__all_sync(B200E_FULL, !stepping || lane_active(L)));
    }
    block_reduce_store<REDUCE>(P, A);
}

// Under the dynamic schedules the static per-node kernel only runs launches with at most one point per lane -- the
// Koopman refinement rounds, where a round costs one trajectory LATENCY, not throughput: it is compiled for half the
// residency (twice the registers: no spills, a freer instruction schedule; lotka_volterra at tol 1e-8, 29 rounds:
// 6.34 -> 5.57 ms of kernel time, profiles/r2_koopman_shapes.log).  With B200E_SCHED_STATIC it carries the large
// batches as well and keeps the full residency.
#if B200E_DYNAMIC
#define B200E_NODES_MINBLOCKS ((B200E_MINBLOCKS + 1) / 2)
#else
#define B200E_NODES_MINBLOCKS B200E_MINBLOCKS
#endif
extern "C" __global__ void __launch_bounds__(B200E_BLOCK, B200E_NODES_MINBLOCKS)
b200e_nodes(const __grid_constant__ b200e_kparams P) {
    tsit5_run<false, false, false>(P);
}
extern "C" __global__ void __launch_bounds__(B200E_BLOCK, B200E_MINBLOCKS)
b200e_reduce(const __grid_constant__ b200e_kparams P) {
    tsit5_run<true, false, false>(P);
}
#if B200E_DYNAMIC
extern "C" __global__ void __launch_bounds__(B200E_BLOCK, B200E_MINBLOCKS)
b200e_nodes_dyn(const __grid_constant__ b200e_kparams P) {
    tsit5_run<false, true, false>(P);
}
extern "C" __global__ void __launch_bounds__(B200E_BLOCK, B200E_MINBLOCKS)
b200e_reduce_dyn(const __grid_constant__ b200e_kparams P) {
    tsit5_run<true, true, false>(P);
}
#endif
// MonteCarlo with the sampler in the kernel (b200e_reduce_generated)
extern "C" __global__ void __launch_bounds__(B200E_BLOCK, B200E_MINBLOCKS)
b200e_reduce_gen(const __grid_constant__ b200e_kparams P) {
    tsit5_run<true, B200E_DYNAMIC != 0, true>(P);
}

#else
// ------------------------------------------------------------------------------------------
// Fixed-step Euler-Maruyama driven by the KKL sine series (process-noise path):
//   u += f(u,p,t) dt + g(u,p,t) dW,  dW = W(t+dt) - W(t),  W(t) = sum_k Z_k phi_k(t)
// phi_k(t_n) does not depend on the sample: the host tabulates it once per handle.
// Every lane takes the same number of steps, so the warp never diverges.
// ------------------------------------------------------------------------------------------
constexpr int NK = B200E_NKKL;

// SPT samples per thread.  Every lane of a warp reads the SAME table row each step, and the L1 returns it
// to all 32 lanes: 64 B x 32 lanes per step saturate the LSU write-back path (ncu: l1tex lsu_writeback_active
// 90 %, long-scoreboard stalls 55 %, FP64 pipe 52 %) long before the FP64 pipe.  A lane that integrates two
// samples side by side uses each loaded row twice: half the load traffic per sample, and the step becomes
// FP64-bound again (measured 3.96 -> 2.57 ms on C5 with two samples per thread at two blocks of 256, 2.36 ms
// with three at three blocks of 128: the host's default).
#ifndef B200E_EM_SPT
#define B200E_EM_SPT 2
#endif
#ifndef B200E_EM_PREFETCH
#define B200E_EM_PREFETCH 1
#endif
constexpr int SPT = B200E_EM_SPT;

template <bool REDUCE, bool DYN, bool GEN>
__device__ __forceinline__ void em_run(const b200e_kparams &P) {
    const long long T = (long long)gridDim.x * B200E_BLOCK;
    const int lane = threadIdx.x & 31;
    Accum A;
    A.init();
    const long long nsteps = P.em_nsteps;
    const real t0 = (real)P.t0, t1 = (real)P.t1, h = (real)P.dt_fixed;
    // a warp works on pools of 32 * SPT consecutive points: lane l takes points base + 32 q + l, q < SPT.
    // Static: pool `warp index + round * warps` ; dynamic: the next pool off the launch's counter (every
    // lane takes the same number of steps, but resident warps do not advance at equal pace).  The host
    // launches the static instantiation when there is at most one point per lane.
    constexpr bool dyn = DYN;
    // Reproducible dynamic schedule: the warp takes a super-chunk of P.super_chunks x 32 consecutive points
    // (a whole number of pools), sums every pool over its lanes with a fixed shuffle tree, adds the pool sums
    // in order and writes one row of sums per super-chunk (see tsit5_run).
    constexpr bool repro = B200E_REPRO && REDUCE && DYN;
    __shared__ double sh_wacc[repro ? NWARPS : 1][repro ? NACC : 1];
    const int pools_per_grab = repro ? (int)(P.super_chunks / SPT) : 1;
    if (repro) {
        for (int k = lane; k < NACC; k += 32) sh_wacc[threadIdx.x >> 5][k] = 0.0;
        __syncwarp();
    }
    for (long long pool = ((long long)blockIdx.x * B200E_BLOCK + threadIdx.x) >> 5;; pool += T >> 5) {
        long long base0;
        if (dyn) {
            unsigned long long b = 0ull;
            if (lane == 0) b = atomicAdd(P.work_counter, (unsigned long long)(32 * SPT) * (unsigned long long)pools_per_grab);
            base0 = (long long)__shfl_sync(B200E_FULL, b, 0);
        } else {
            base0 = pool * (32 * SPT);
        }
        if (base0 >= P.npts) break;
      for (int pg = 0; pg < pools_per_grab; pg++) {
        const long long base = base0 + (long long)pg * (32 * SPT);
        if (base >= P.npts) break;

        long long idx[SPT];
        bool valid[SPT];
        double wgt[SPT];
        real u[SPT][N], p[SPT][B200E_DIM(NP)], z[SPT][NK];
#pragma unroll
        for (int q = 0; q < SPT; q++) {
            idx[q] = base + 32 * q + lane;
            valid[q] = idx[q] < P.npts;
            const long long il = valid[q] ? idx[q] : P.npts - 1;  // ragged last pool: solve a copy, emit nothing
            double xs[NX];
            load_point<GEN>(P, il, xs);
            wgt[q] = P.weight_by_pdf ? pdf_weight(P, xs) : 1.0;
            real xr[NX], u0n[N], pn[B200E_DIM(NP)];
#pragma unroll
            for (int j = 0; j < NX; j++) xr[j] = (real)xs[j];
#pragma unroll
            for (int j = 0; j < N; j++) { u0n[j] = (real)P.u0_nom[j]; u[q][j] = u0n[j]; }
#pragma unroll
            for (int j = 0; j < NP; j++) { pn[j] = (real)P.p_nom[j]; p[q][j] = pn[j]; }
#pragma unroll
            for (int k = 0; k < NK; k++) z[q][k] = (real)0;
            // Z, p = h(x, u0, p); u0 stays nominal (expectation.jl:216)
            hmap(xr, u0n, pn, z[q], p[q]);
        }

        // W(t_n) = sum_k Z_k phi_k(t_n) from the table row of step n (one 16-byte load per pair of modes,
        // shared by the lane's SPT samples).  Step times as in the oracle, t_n = t0 + n h with t_N = t1, from
        // a floating-point counter (an int64 -> double conversion per step costs more than the whole drift).
        const double *tab = P.kkl_table;
        real wcur[SPT];
        auto load_row = [&](real *row, long long idx) {
            const double *r = tab + idx * NK;
            if (NK % 2 == 0) {
                const double2 *r2 = reinterpret_cast<const double2 *>(r);
#pragma unroll
                for (int k = 0; k < NK / 2; k++) {
                    const double2 v = __ldg(r2 + k);
                    row[2 * k] = (real)v.x;
                    row[2 * k + 1] = (real)v.y;
                }
            } else {
#pragma unroll
                for (int k = 0; k < NK; k++) row[k] = (real)__ldg(r + k);
            }
        };
        auto kkl_dot = [&](const real *zq, const real *row) {
            real w = (real)0;
#pragma unroll
            for (int k = 0; k < NK; k++) w = fma(zq[k], row[k], w);
            return w;
        };
        const int ns = (int)nsteps;
        real t = t0, sd = (real)0;
        // one Euler-Maruyama step of the lane's SPT samples with the table row of t_{s+1}
        auto advance = [&](const real *row, int s) {
            sd += (real)1;
            const real tn = (s + 1 == ns) ? t1 : fma(sd, h, t0);
            const real dt = tn - t;
#pragma unroll
            for (int q = 0; q < SPT; q++) {
                const real wn = kkl_dot(z[q], row);
                const real dW = wn - wcur[q];
                wcur[q] = wn;
                real f[N], g[N];
                rhs(f, u[q], p[q], t);
                diffusion(g, u[q], p[q], t);
#pragma unroll
                for (int j = 0; j < N; j++) u[q][j] = fma(g[j], dW, fma(dt, f[j], u[q][j]));
            }
            t = tn;
        };
        real rowA[NK];
        load_row(rowA, 0);
#pragma unroll
        for (int q = 0; q < SPT; q++) wcur[q] = kkl_dot(z[q], rowA);
#if B200E_EM_PREFETCH
        // two-stage software pipeline over the steps: the row of step s + 2 is requested before step s is computed,
        // so its L1 latency is covered by a whole step of arithmetic (the warp count is low here: registers)
        real rowB[NK];
        load_row(rowA, ns < 1 ? 0 : 1);
        int s = 0;
        for (; s + 1 < ns; s += 2) {
            load_row(rowB, s + 2);
            advance(rowA, s);
            load_row(rowA, s + 3 < ns ? s + 3 : ns);
            advance(rowB, s + 1);
        }
        if (s < ns) advance(rowA, s);
#else
        for (int s = 0; s < ns; s++) {
            load_row(rowA, s + 1);
            advance(rowA, s);
        }
#endif
#pragma unroll
        for (int q = 0; q < SPT; q++) {
            if (repro) {  // chunk sum in a fixed order: lanes without a point contribute zeros
                real out[NOUT];
                observable(u[q], p[q], (real)P.t1, out);
#pragma unroll 1
                for (int k = 0; k < NOUT; k++) {
                    double v = valid[q] ? (double)out[k] : 0.0;
                    if (P.weight_by_pdf) v *= wgt[q];
                    double v2 = v * v;
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
                        v += __shfl_xor_sync(B200E_FULL, v, off);
                        v2 += __shfl_xor_sync(B200E_FULL, v2, off);
                    }
                    if (lane == 0) {
                        sh_wacc[threadIdx.x >> 5][k] += v;
                        sh_wacc[threadIdx.x >> 5][NOUT + k] += v2;
                    }
                }
            }
            if (!valid[q]) continue;
            int failed = 0;
#pragma unroll
            for (int j = 0; j < N; j++) failed |= !(u[q][j] == u[q][j]);
            if (!repro) emit<REDUCE>(P, idx[q], u[q], p[q], wgt[q], A);
            if (!REDUCE && P.steps) P.steps[idx[q]] = (int)nsteps;
            A.count((unsigned)nsteps, (unsigned)nsteps, 0u, (unsigned)failed);
        }
      }  // pools of this grab
        if (repro) {  // the super-chunk is complete: its row leaves the warp
            const long long row = base0 / ((long long)(32 * SPT) * pools_per_grab);
            __syncwarp();
            for (int k = lane; k < NACC; k += 32) {
                P.chunk_rows[row * NACC + k] = sh_wacc[threadIdx.x >> 5][k];
                sh_wacc[threadIdx.x >> 5][k] = 0.0;
            }
            __syncwarp();
        }
    }
    block_reduce_store<REDUCE>(P, A);
}

extern "C" __global__ void __launch_bounds__(B200E_BLOCK, B200E_MINBLOCKS)
b200e_nodes(const __grid_constant__ b200e_kparams P) {
    em_run<false, false, false>(P);
}
extern "C" __global__ void __launch_bounds__(B200E_BLOCK, B200E_MINBLOCKS)
b200e_reduce(const __grid_constant__ b200e_kparams P) {
    em_run<true, false, false>(P);
}
#if B200E_DYNAMIC
extern "C" __global__ void __launch_bounds__(B200E_BLOCK, B200E_MINBLOCKS)
b200e_nodes_dyn(const __grid_constant__ b200e_kparams P) {
    em_run<false, true, false>(P);
}
extern "C" __global__ void __launch_bounds__(B200E_BLOCK, B200E_MINBLOCKS)
b200e_reduce_dyn(const __grid_constant__ b200e_kparams P) {
    em_run<true, true, false>(P);
}
#endif
// MonteCarlo with the sampler in the kernel (b200e_reduce_generated)
extern "C" __global__ void __launch_bounds__(B200E_BLOCK, B200E_MINBLOCKS)
b200e_reduce_gen(const __grid_constant__ b200e_kparams P) {
    em_run<true, B200E_DYNAMIC != 0, true>(P);
}
#endif

}  // namespace b200e
