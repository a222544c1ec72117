/*
 * b200_expect.h -- C ABI of libb200expect.so: the B200-native drop-in for the batched
 * expectation integrand of SciMLExpectations.jl.
 *
 * Reference interfaces each entry point replaces (paths relative to the reference repo):
 *
 *   b200e_create          <- SystemMap(prob, alg, ensemblealg; kwargs...)     src/system_utils.jl:32-53
 *                            + ExpectationProblem(sm, g, h, d)                src/problem_types.jl:60-66, :81-85
 *                            + GenericDistribution(dists...)                  src/distribution_utils.jl:40-48
 *                            (the data the hot path closes over: template problem, solver and its
 *                             kwargs, covariate map h, observable g, product density d)
 *   b200e_eval_nodes      <- integrand_koopman_systemmap_batch!(dx, x, p)     src/expectation.jl:182-196
 *                            and its process-noise twin                       src/expectation.jl:229-243
 *                            (the f!(dx, x, p) Integrals.jl calls once per cubature refinement round)
 *   b200e_reduce_samples  <- solve(exprob, MonteCarlo(n)): the EnsembleProblem solve + mean(sol.u)
 *                                                                             src/expectation.jl:277-293, :304-324
 *   b200e_reduce_samples_begin / _end
 *                         <- the same reduction in two halves (device-resident points): a caller's own exchange
 *                            step overlaps the next kernel
 *   b200e_reduce_generated <- the same solve with rand(d) of prob_func (src/expectation.jl:278,
 *                            src/distribution_utils.jl:43) drawn inside the kernel from a counter-based stream;
 *                            b200e_sample_host / b200e_sample_device reproduce that stream bit for bit
 *   b200e_eval_regions    <- the Genz-Malik rule Cubature's hcubature_v applies to the dx it gets back from
 *                            the integrand callback (third-party C library behind src/expectation.jl:361-362;
 *                            SURVEY.md 8f rank 1): regions in, (value, error, split dimension) per region out
 *   b200e_koopman_solve   <- solve(exprob, Koopman(); quadalg=CubatureJLh(), batch=...)  src/expectation.jl:336-354
 *                            with hcubature_v's region heap kept inside the library
 *
 * Conventions: every function returns 0 on success and a negative B200E_ERR_* code on failure;
 * b200e_last_error(h) returns a message owned by the library.  All host buffers are caller-owned;
 * the library copies in/out inside the call and keeps no caller pointer after it returns (the one exception is
 * spelled out at b200e_reduce_samples_begin).  Calls on
 * one handle are serialised by the caller; distinct handles may be used from distinct threads.
 * x / dx / sum pointers may be HOST pointers (pageable or pinned) or DEVICE pointers on the
 * handle's device (detected with cudaPointerGetAttributes); device pointers skip the copies.
 * There is NO CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef B200_EXPECT_H
#define B200_EXPECT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200E_ABI_VERSION 4

/* limits (compile-time dims of the generated kernel) */
#define B200E_MAX_STATE 16
#define B200E_MAX_PARAM 32
#define B200E_MAX_X 32
#define B200E_MAX_OUT 64
#define B200E_MAX_SAVE 64
#define B200E_MAX_KKL 32

enum {
    B200E_OK = 0,
    B200E_ERR_INVALID = -1,   /* bad argument / dimension */
    B200E_ERR_COMPILE = -2,   /* NVRTC failed; see b200e_compile_log */
    B200E_ERR_CUDA = -3,      /* CUDA runtime/driver error */
    B200E_ERR_NO_DEVICE = -4, /* no usable CUDA device (there is no CPU fallback) */
    B200E_ERR_NOMEM = -5,
    B200E_ERR_NCCL = -6
};

/* TSIT5: adaptive Tsit5 (the ODE path, src/expectation.jl:191 / :290).  EM: fixed-step Euler-Maruyama
 * driven by the KKL noise path (src/expectation.jl:215-225 / :304-314 with dt = dt_fixed).  LAMBA_EM: the
 * same path with StochasticDiffEq's adaptive `LambaEM`, the solver the reference's own process-noise test
 * uses (test/processnoise.jl:15: ProcessNoiseSystemMap(prob, 8, LambaEM(), abstol, reltol)); tolerances
 * from abstol / reltol, dt_fixed unused. */
enum { B200E_SOLVER_TSIT5 = 0, B200E_SOLVER_EM = 1, B200E_SOLVER_LAMBA_EM = 2 };
enum { B200E_FP64 = 0, B200E_FP32 = 1 };
/* layout of x (and dx): 0 = point-major x[i*nx+j] (what Julia's dim x npts column-major array is),
 *                        1 = SoA x[j*npts+i] */
enum { B200E_LAYOUT_POINT_MAJOR = 0, B200E_LAYOUT_SOA = 1 };
/* how points are handed to warps.  DYNAMIC (default): a warp whose lanes run idle takes the next
 * unassigned points of the launch off a device-side counter, so no warp waits for work another one
 * still owns (7-18 % faster than STATIC: the SM's warp scheduler does not advance resident warps at
 * equal pace).  Per-node outputs and all integer counters are identical under both; the floating-point
 * SUMS of b200e_reduce_samples are accumulated per lane, so under DYNAMIC their last bits depend on the
 * run-time assignment (differences ~1e-16 relative).  STATIC: lane l owns points l, l+T, l+2T, ...;
 * sums are bit-reproducible for a given launch shape.  DYNAMIC_REPRODUCIBLE (end-state and time-sampled observables):
 * warps take super-chunks of consecutive points, sum each chunk of 32 with a fixed shuffle tree and
 * write one row of sums per super-chunk; the rows are folded in index order.  Sums are bit-reproducible and do
 * not depend on the launch shape either, at the price of always starting 32 trajectories together. */
enum { B200E_SCHED_DYNAMIC = 0, B200E_SCHED_STATIC = 1, B200E_SCHED_DYNAMIC_REPRODUCIBLE = 2 };
enum { B200E_PDF_NONE = 0, B200E_PDF_UNIFORM = 1, B200E_PDF_NORMAL = 2, B200E_PDF_TRUNCNORMAL = 3, B200E_PDF_TABLE = 4,
       B200E_PDF_LOGPOLY = 5 };

/* one factor of GenericDistribution(dists...) (src/distribution_utils.jl:40-48) */
typedef struct {
    int32_t kind; /* B200E_PDF_* */
    int32_t _pad;
    double lo, hi;    /* support: uniform bounds / truncation bounds (+-inf for plain normal) / table grid ends */
    double mu, sigma; /* normal parameters */
    /* B200E_PDF_TABLE: density VALUES on the uniform grid lo + i (hi - lo) / (ntable - 1) -- `k.density` / `k.x` of a
     * KernelDensity.UnivariateKDE -- evaluated as pdf(kde, x) does: the quadratic B-spline through the values
     * (Interpolations' BSpline(Quadratic(Line(OnGrid()))): C1, zero second derivative in the end cells), 0 outside
     * [lo, hi]  (docs/src/tutorials/gpu_bayesian.md:153-166: pdf_func(x) = prod(pdf.(p_kde, x)), lb / ub = the grid
     * ends).  b200e_create prefilters the values into spline coefficients on the host and copies those; the kernel
     * evaluates three taps.  Koopman only: a tabulated factor cannot be sampled. */
    /* B200E_PDF_LOGPOLY: any factor whose log-density on [lo, hi] is
     *     logpdf(x) = c0 + c1 x + c2 x^2 + c3 log(x) + c4 log(1 - x) + c5 log(x)^2          (-inf outside [lo, hi])
     * -- Exponential, Gamma / Chi-squared / Erlang, Beta, LogNormal, Rayleigh, Normal, and truncated(...) of any of them
     * with log(cdf(hi) - cdf(lo)) folded into c0 by the caller (Distributions.logpdf semantics behind
     * src/distribution_utils.jl:42).  `table` points to the six coefficients c0 .. c5, `ntable` = 6; lo / hi may be
     * infinite (a MonteCarlo model; a Koopman solve needs a finite box).  Like a tabulated factor it cannot be drawn by
     * the in-kernel sampler: MonteCarlo ships host samples for such a model. */
    const double *table;
    int64_t ntable;
} b200e_pdf_dim;

/*
 * Everything the hot path closes over.  `source` holds CUDA device functions, compiled by NVRTC
 * for sm_100a into the generated kernel (arbitrary Julia closures cannot cross a C ABI):
 *
 *   B200E_FN void rhs(real* du, const real* u, const real* p, real t);                (ODE f / SDE drift)
 *   B200E_FN void hmap(const real* x, const real* u0_nom, const real* p_nom,
 *                      real* u0_or_Z, real* p);                 (h, expectation.jl:176, :216)
 *   B200E_FN void observable(const real* u_end, const real* p, real t_end, real* out); (g, end-state form)
 *   B200E_FN void diffusion(real* g, const real* u, const real* p, real t);           (EM only; scalar noise)
 *   B200E_FN void observable_at(int k, real t_k, const real* u_k, const real* p, real* out_k);
 *                      (g built from the solution at saved times -- sol(t_k), sol[i, :] with saveat,
 *                       docs/src/tutorials/introduction.md:53, :192, :207-212: called once per save time, in
 *                       order, with the state interpolated by Tsit5's free 4th-order interpolant; writes the
 *                       nout / nsave outputs of slot k.  Replaces `observable` when nsave > 0.)
 *
 * u0/p outputs of hmap are pre-filled with the nominal values, so hmap only writes what it changes.
 */
typedef struct {
    uint32_t struct_size; /* sizeof(b200e_model), for ABI evolution */
    int32_t nstate, nparam, nx, nout;
    int32_t solver;  /* B200E_SOLVER_* */
    int32_t n_kkl;   /* EM / LAMBA_EM: number of KKL terms (length of Z); 0 for Tsit5 */
    int32_t fp_mode; /* B200E_FP64 / B200E_FP32 (state arithmetic; accumulation is always f64) */
    const char *source;
    const double *u0; /* nominal initial condition, nstate entries (prob.u0) */
    const double *p;  /* nominal parameters, nparam entries (prob.p) */
    double t0, t1;    /* prob.tspan */
    double abstol, reltol; /* Tsit5 / LambaEM tolerances (OrdinaryDiffEq defaults 1e-6 / 1e-3) */
    double dtmax;          /* <= 0 -> t1 - t0 */
    double dt_fixed;       /* EM step */
    int64_t maxiters;      /* <= 0 -> 100000 */
    const b200e_pdf_dim *pdf; /* nx entries, or NULL for weight 1 */
    /* placement and launch shape (0 = library default) */
    int32_t n_devices;        /* 0/1: single device */
    const int32_t *devices;   /* CUDA ordinals; NULL -> current device (n_devices<=1) or 0..n-1 */
    int32_t block_size;       /* threads per block, multiple of 32 (default 256) */
    int32_t blocks_per_sm;    /* resident blocks per SM to launch (grid = SMs * this) */
    int32_t refill_threshold; /* lanes of a warp that must be idle before the warp reloads (1..32); 0 = automatic */
    int32_t chunk_points;     /* host-pointer calls are pipelined in chunks of this many points */
    int32_t schedule;         /* B200E_SCHED_* */
    int32_t nsave;            /* > 0: time-sampled observable (Tsit5 only), nout = nsave * outputs per save time */
    const double *save_times; /* nsave ascending times in [t0, t1] (saveat / the arguments of sol(t)) */
    const char *cache_dir;    /* cubin cache directory, NULL -> $B200E_CACHE_DIR or disabled */
} b200e_model;

typedef struct {
    uint64_t nf;      /* RHS evaluations */
    uint64_t naccept; /* accepted steps */
    uint64_t nreject; /* rejected steps */
    uint64_t nfail;   /* trajectories aborted (maxiters / non-finite / dt underflow) */
} b200e_counters;

/* timing/shape of the most recent call on a handle (filled from CUDA events on the launch stream) */
typedef struct {
    double kernel_ms;   /* sum of main-kernel durations of the last call (device 0 of the handle) */
    double total_ms;    /* first enqueue -> last completion, CUDA events */
    int64_t launches;   /* kernel launches (main + finalize) in the last call */
    int64_t h2d_bytes, d2h_bytes;
    int32_t grid, block, regs_per_thread, blocks_per_sm_occ;
    int32_t local_bytes_per_thread, smem_bytes;
} b200e_call_stats;

typedef struct b200e_handle b200e_handle;

int b200e_abi_version(void);

/* NVRTC-compile the model for sm_100a (cached by content hash) and set up device state. */
int b200e_create(const b200e_model *model, b200e_handle **out);
int b200e_destroy(b200e_handle *h);

/* Compile only (needs no GPU): returns the cubin size and leaves the ptxas/NVRTC log in
 * b200e_last_error(NULL).  Used by build() to cross-check models for sm_100a on a CPU box. */
int b200e_compile_check(const b200e_model *model, int64_t *cubin_bytes);

/* message of the last failure on this handle (h == NULL: last failure of create/compile_check
 * on the calling thread).  Never NULL. */
const char *b200e_last_error(const b200e_handle *h);
/* NVRTC + ptxas -v log of the handle's kernel */
const char *b200e_compile_log(const b200e_handle *h);

/* Koopman batch integrand: dx[:, i] = g(sol_i, p_i) * pdf(d, x[:, i])  (weight_by_pdf != 0)
 * Blocking: dx is completely written when the call returns. */
int b200e_eval_nodes(b200e_handle *h, const double *x, int64_t npts, int32_t layout,
                     int32_t weight_by_pdf, double *dx);

/* MonteCarlo block: sum[k] = sum_i g_k(sol_i, p_i) [* pdf], sumsq[k] = sum_i (.)^2 (may be NULL),
 * counters may be NULL.  The caller divides by n (mean(sol.u), expectation.jl:293).
 * With n_devices > 1 the samples are sharded in contiguous index ranges and the per-device
 * partial sums are combined by one NCCL all-reduce. */
int b200e_reduce_samples(b200e_handle *h, const double *x, int64_t n, int32_t layout,
                         int32_t weight_by_pdf, double *sum, double *sumsq, b200e_counters *c);

/* The same reduction split in two for device-resident points on a single-device handle: _begin enqueues the
 * kernel (and the copy of its 2*nout + 4 results) and returns, _end waits and hands the results over.  A caller
 * with an exchange step of its own -- one process per GPU, the all-reduce of the previous batch -- puts that step
 * between the two and overlaps it with the kernel.  Up to TWO reductions may be pending per handle (first in, first
 * out): with the next batch queued behind the running one the device never waits for the host between batches
 * (a streamed MonteCarlo(10^9), block after block).  Each x must stay valid and unchanged until its _end returns;
 * no other call on the handle is accepted while a reduction is pending. */
int b200e_reduce_samples_begin(b200e_handle *h, const double *x, int64_t n, int32_t layout, int32_t weight_by_pdf);
int b200e_reduce_samples_end(b200e_handle *h, double *sum, double *sumsq, b200e_counters *c);

/* counters of the last b200e_eval_nodes call */
int b200e_last_counters(const b200e_handle *h, b200e_counters *c);
int b200e_last_stats(const b200e_handle *h, b200e_call_stats *s);

/* optional per-point diagnostics of the next eval_nodes call: attempted steps per point
 * (host int32 buffer of npts entries; NULL disables) */
int b200e_set_steps_buffer(b200e_handle *h, int32_t *steps);

/* pinned host memory for callers that want true async H2D (cudaHostAlloc / cudaFreeHost) */
int b200e_host_alloc(void **ptr, int64_t bytes);
int b200e_host_free(void *ptr);
/* device memory on the handle's (first) device, for keeping sample sets resident in HBM */
int b200e_device_alloc(b200e_handle *h, void **dptr, int64_t bytes);
int b200e_device_free(b200e_handle *h, void *dptr);
int b200e_memcpy_h2d(b200e_handle *h, void *dptr, const void *src, int64_t bytes);
int b200e_memcpy_d2h(b200e_handle *h, void *dst, const void *dptr, int64_t bytes);

/* ---- MonteCarlo with device-side sampling --------------------------------------
 * Replaces `rand(d)` inside prob_func of solve(exprob, MonteCarlo(n)) (src/expectation.jl:277-280, :304-305;
 * src/distribution_utils.jl:43).  Sample i of stream `seed` is a pure function of (seed, i): Philox4x32-10
 * -> 52-bit uniforms -> inverse CDF of each factor of the model's product density, written with exactly
 * rounded operations only (csrc/kernels/expect_sampling.cuh), so the three entry points below produce
 * bit-identical doubles on the host and on the device. */

/* sum / sumsq / counters over samples first_index .. first_index + n - 1 of stream `seed`, drawn inside
 * the kernel (weight 1, as MonteCarlo): nothing but the 2*nout + 4 results crosses PCIe. */
int b200e_reduce_generated(b200e_handle *h, uint64_t seed, uint64_t first_index, int64_t n, double *sum,
                           double *sumsq, b200e_counters *c);
/* the host twin: the same samples into caller memory, for the CPU arm and for tests.  Takes the model, not a
 * handle: it needs no device (errors go to b200e_last_error(NULL)). */
int b200e_sample_host(const b200e_model *model, uint64_t seed, uint64_t first_index, int64_t n, int32_t layout,
                      double *x);
/* the same samples written by the device (x: host or device pointer) */
int b200e_sample_device(b200e_handle *h, uint64_t seed, uint64_t first_index, int64_t n, int32_t layout, double *x);

/* ---- the caller above the path, moved next to it (optional fast path for Koopman) ---------------- */
typedef struct {
    double reltol, abstol; /* ireltol / iabstol of solve(exprob, Koopman(); ...) */
    int64_t maxevals;      /* maxiters of the reference call; <= 0: unlimited */
} b200e_cubature_opts;

typedef struct {
    int64_t nevals, nrounds, nregions, max_batch;
    int32_t converged, _pad;
    double kernel_ms, total_ms; /* device time of all rounds (CUDA events) / wall time of the solve */
    b200e_counters counters;    /* summed over all rounds */
} b200e_cubature_stats;

/* One refinement round: for each region (centre c, half-width w; nx doubles each, region-major) the
 * device generates the nodes of the Genz-Malik 7/5 rule (Gauss-Kronrod 7/15 for nx == 1), evaluates
 * g(sol,p)*pdf at each, and returns per region val[nout], err[nout] = |I7 - I5| and the split
 * dimension (largest summed fourth difference).  Host pointers; blocking. */
int b200e_eval_regions(b200e_handle *h, const double *centers, const double *halfwidths, int64_t nregions,
                       double *val, double *err, int32_t *splitdim);

/* Whole h-adaptive solve over the box [lb, ub] (error_norm INDIVIDUAL): u[nout], resid[nout]. */
int b200e_koopman_solve(b200e_handle *h, const double *lb, const double *ub, const b200e_cubature_opts *opts,
                        double *u, double *resid, b200e_cubature_stats *stats);

/* FP64 (fp32 != 0: FP32) FMA-chain microbenchmark on the handle's device: the measured
 * roofline denominator for this path (MEASURED_PEAKS.json has no FP64 entry). */
int b200e_measure_fma_peak(int32_t device, int32_t fp32, double *tflops);

/* number of visible CUDA devices (0 when there is none; never fails) */
int b200e_device_count(void);

#ifdef __cplusplus
}
#endif
#endif
