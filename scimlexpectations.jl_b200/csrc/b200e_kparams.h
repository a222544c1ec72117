// b200e_kparams.h -- launch parameters shared verbatim by the host library (nvcc) and the
// NVRTC-generated kernels.  Plain C types only (NVRTC has no <stdint.h>); every field is naturally
// aligned so host and device layouts agree.
#ifndef B200E_KPARAMS_H
#define B200E_KPARAMS_H

#define B200E_KP_MAX_STATE 16
#define B200E_KP_MAX_PARAM 32
#define B200E_KP_MAX_X 32
#define B200E_KP_MAX_SAVE 64

// device form of one factor of the product density: logpdf(x) = logc - 0.5*((x-mu)*inv_sigma)^2
// inside [lo,hi], -inf outside.  logc folds -log(sigma) - log(2pi)/2 - log(cdf(hi)-cdf(lo))
// (or -log(hi-lo) for uniform, where inv_sigma = 0) and is computed once on the host.
// smp_*: constants of the counter-based sampler (kernels/expect_sampling.cuh): uniform (lo, hi-lo, -, -),
// normal / truncated normal (mu, sigma, cdf(lo), cdf(hi) - cdf(lo)).
struct b200e_kpdf {
    double lo, hi, mu, inv_sigma, logc;
    double smp_a, smp_b, smp_c, smp_d;
    const double *tab;   // kind 4: ntab + 2 quadratic B-spline coefficients of the density on the uniform grid lo + i / inv_dx (device memory; knot i <-> tab[i + 1])
    double inv_dx;
    int kind;
    int ntab;
};

struct b200e_kparams {
    const double *x;        // sample points / cubature nodes (device)
    double *dx;             // per-node outputs (nodes mode)
    double *partials;       // [grid][2*nout] per-block partial sums (reduce mode)
    unsigned long long *cpartials;  // [grid][4] per-block counters
    int *steps;             // optional per-point attempted steps (nodes mode), or null
    unsigned long long *work_counter;  // [0] next unassigned point of this launch (dynamic work distribution),
                                       // [1] blocks that have stored their partials; both are zero between
                                       // launches: the last block to finish resets them
    double *final_sums;                // optional [2*nout]: the last block folds all per-block partials of the launch
    unsigned long long *final_counts;  // optional [4]: same for the counters
    double *chunk_rows;                // reproducible dynamic schedule: [rows][2*nout] sums, one row per super-chunk
    long long super_chunks;            //   chunks of 32 consecutive points per super-chunk
    const double *kkl_table;  // EM: [nsteps+1][n_kkl] basis values sqrt(2) sin((k-1/2) pi t_n)/((k-1/2) pi)
    long long npts;         // points in this launch
    long long ld;           // SoA leading dimension of x and dx (total points of the caller's array)
    long long maxiters;
    long long em_nsteps;
    double t0, t1;
    double abstol, reltol, dtmax, dt_fixed;
    double snap_tol;        // 100 * ulp(|t1|): end-point snapping of the last step
    int layout;             // 0 point-major, 1 SoA
    int weight_by_pdf;
    int refill_threshold;
    int want_sumsq;
    int gen;                // 1: points are drawn in the kernel (sample gen_first + i of stream gen_seed), x is unused
    int _pad0;
    unsigned long long gen_seed;
    long long gen_first;
    double u0_nom[B200E_KP_MAX_STATE];
    double p_nom[B200E_KP_MAX_PARAM];
    b200e_kpdf pdf[B200E_KP_MAX_X];
    double save_t[B200E_KP_MAX_SAVE];  // time-sampled observables: ascending save times (B200E_NSAVE of them)
};

#endif
