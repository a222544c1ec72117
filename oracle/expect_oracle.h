/*
 * expect_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY).
 *
 * A plain-C restatement of the batched expectation integrand of
 * SciMLExpectations.jl (reference src/expectation.jl:169-200, :268-294,
 * :209-249, :295-325) and of the third-party numerics that path delegates to
 * (OrdinaryDiffEq Tsit5 + PI controller, Distributions logpdf, the KKL noise
 * series, Euler-Maruyama).  The reference is Julia and Julia is absent from
 * this image, so this oracle is a "port": every function cites the reference
 * file:line (or the named dependency) it follows.
 *
 * PARITY PIN: the README golden vector (reference README.md:65-69) pins the
 * Tsit5/controller restatement to ~2e-7 through tests/test_oracle_golden.py;
 * the reference holds no fixture tighter than rtol=1e-2, so at 1e-8 the
 * contract is kernel-vs-oracle on identical inputs.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product never does.
 */
#ifndef EXPECT_ORACLE_H
#define EXPECT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_STATE 16
#define ORC_MAX_PARAM 32
#define ORC_MAX_X 32
#define ORC_MAX_OUT 64
#define ORC_MAX_SAVE 64
#define ORC_MAX_KKL 32

typedef void (*orc_rhs_fn)(double *du, const double *u, const double *p, double t);
/* h(x, u0_nominal, p_nominal) -> (u0, p); for the process-noise path the first
 * output is Z (the KKL coefficients), reference expectation.jl:216 / :305. */
typedef void (*orc_map_fn)(const double *x, const double *u0_nom, const double *p_nom,
                           double *u0_or_z, double *p);
/* g(sol, p) restricted to end-state observables: g(u(t1), p, t1) -> out[nout] */
typedef void (*orc_obs_fn)(const double *u_end, const double *p, double t_end, double *out);

/* g(sol, p) built from the solution at saved times (sol(t_k), sol[i, :] with saveat; reference
 * docs/src/tutorials/introduction.md:53, :192, :207-212): called once per save time, in order, with the
 * interpolated state; writes the nout / nsave outputs of slot k */
typedef void (*orc_obs_at_fn)(int k, double t_k, const double *u_k, const double *p, double *out_k);

enum { ORC_PDF_NONE = 0, ORC_PDF_UNIFORM = 1, ORC_PDF_NORMAL = 2, ORC_PDF_TRUNCNORMAL = 3, ORC_PDF_TABLE = 4,
       ORC_PDF_LOGPOLY = 5 /* logpdf = c0 + c1 x + c2 x^2 + c3 log x + c4 log(1-x) + c5 log(x)^2 on [lo, hi]; table = c0..c5 */ };

/* one factor of GenericDistribution(dists...), reference distribution_utils.jl:40-48 */
typedef struct {
    int32_t kind;
    int32_t _pad;
    double lo, hi;    /* support (uniform bounds / truncation bounds); +-inf for normal */
    double mu, sigma; /* normal parameters (unused for uniform) */
    /* ORC_PDF_TABLE: pdf(kde, x) of a KernelDensity estimate (docs/src/tutorials/gpu_bayesian.md:153-166) on the
     * uniform grid lo .. hi of ntable points: quadratic B-spline (Interpolations' Quadratic(Line(OnGrid()))), 0
     * outside.  `table` holds the ntable + 2 prefiltered coefficients orc_kde_prefilter makes of the grid values */
    const double *table;
    int64_t ntable;
} orc_pdf_dim;

enum { ORC_SOLVER_TSIT5 = 0, ORC_SOLVER_EM = 1, ORC_SOLVER_LAMBA_EM = 2 };

typedef struct {
    int32_t nstate, nparam, nx, nout;
    int32_t solver;   /* ORC_SOLVER_* */
    int32_t n_kkl;    /* EM / LambaEM: number of KKL terms (= length of Z) */
    orc_rhs_fn rhs;
    orc_rhs_fn diffusion; /* EM / LambaEM: g(du,u,p,t), scalar noise multiplies every component */
    orc_map_fn hmap;
    orc_obs_fn obs;
    const double *u0_nom;
    const double *p_nom;
    double t0, t1;
    double abstol, reltol; /* Tsit5 (OrdinaryDiffEq defaults 1e-6 / 1e-3) and LambaEM */
    double dtmax;          /* <=0 -> t1-t0 */
    double dt_fixed;       /* EM step */
    int64_t maxiters;      /* <=0 -> 100000 */
    const orc_pdf_dim *pdf; /* nx entries, or NULL (weight 1) */
    /* time-sampled observables (Tsit5 only): nsave > 0 -> nout = nsave * (outputs per save time), the
     * outputs come from obs_at at the ascending times save_t[0..nsave) in [t0, t1] (obs is not called).
     * States between steps come from Tsit5's free 4th-order interpolant, as OrdinaryDiffEq's saveat /
     * sol(t) do; a save time equal to the end of a step takes the step's own end state. */
    int32_t nsave;
    int32_t _pad;
    const double *save_t;
    orc_obs_at_fn obs_at;
} orc_problem;

typedef struct {
    uint64_t nf;      /* RHS evaluations */
    uint64_t naccept; /* accepted steps */
    uint64_t nreject; /* rejected steps */
    uint64_t nfail;   /* trajectories that aborted (maxiters / non-finite / dt underflow) */
} orc_counters;

/* log pdf of the product distribution at x (sum of per-dim logpdf); -inf outside support */
double orc_logpdf(const orc_pdf_dim *pdf, int nx, const double *x);
/* pdf of the product distribution at x: exp(sum of logpdf) times the tabulated (KDE) factors */
double orc_pdf(const orc_pdf_dim *pdf, int nx, const double *x);
/* grid values f[n] of a KDE -> the n + 2 quadratic B-spline coefficients pdf(kde, x) evaluates */
void orc_kde_prefilter(const double *f, int64_t n, double *c);

/* one trajectory: x -> h -> solve -> g (-> * pdf).  Returns 0 ok, 1 if the trajectory failed
 * (g is still applied to the last state, as the reference does with failed retcodes). */
int orc_solve_point(const orc_problem *pb, const double *x, int weight_by_pdf, double *out,
                    orc_counters *c);

/* Koopman batch integrand, reference expectation.jl:182-196.
 * layout 0: point-major x[i*nx+j], dx[i*nout+k]  (Julia dim x npts column-major)
 * layout 1: SoA         x[j*npts+i], dx[k*npts+i]
 * per_point_steps (optional, npts entries): attempted steps of trajectory i. */
int orc_eval_nodes(const orc_problem *pb, const double *x, int64_t npts, int layout,
                   int weight_by_pdf, double *dx, orc_counters *c, int32_t *per_point_steps,
                   int nthreads);

/* MonteCarlo reduction, reference expectation.jl:277-293: sum_i g_i (and sum_i g_i^2). The
 * caller divides by n.  Deterministic: fixed 4096-sample chunks, Neumaier-compensated. */
int orc_reduce_samples(const orc_problem *pb, const double *x, int64_t n, int layout,
                       int weight_by_pdf, double *sum, double *sumsq, orc_counters *c,
                       int nthreads);

/* KKL series W(t) = sqrt(2) sum_k Z_k sin((k-1/2) pi t)/((k-1/2) pi), expectation.jl:217-220 */
double orc_kkl_w(const double *z, int n, double t);

/* built-in problems for BASELINE.json configs (models_builtin.c).  name in
 * {"linear2","lotka_volterra","lorenz63","decay1","oscillator_noise","gbm4"} */
int orc_builtin_problem(const char *name, orc_problem *out);

int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
