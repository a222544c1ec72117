/*
 * models_builtin.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY).
 *
 * The BASELINE.json configs as plain C callables, written independently of the product's CUDA
 * model sources (scimlexpectations.jl_b200/models.py) so that a transcription error on either
 * side shows up as a parity failure.
 */
#include "expect_oracle.h"

#include <math.h>
#include <string.h>

/* ---------- C1/C2: README linear ODE (reference README.md:31-46, test/solve.jl:17-34) ------ */
/* A = [0 1; -p1 -p2] is built ONCE from the nominal p=[1,2] and captured by the closure
 * (README.md:38-39), so it does not follow the remade p. du .= A*u */
static void linear2_rhs(double *du, const double *u, const double *p, double t) {
    (void)p; (void)t;
    du[0] = 0.0 * u[0] + 1.0 * u[1];
    du[1] = -1.0 * u[0] + -2.0 * u[1];
}
/* cov(x,u,p) = x, p  (README.md:42) */
static void map_x_to_u0(const double *x, const double *u0n, const double *pn, double *u0, double *p) {
    (void)u0n; (void)pn; (void)p;
    u0[0] = x[0];
    u0[1] = x[1];
}
/* g(sol,p) = sol[:,end]  (README.md:58) */
static void obs_state2(const double *u, const double *p, double t, double *out) {
    (void)p; (void)t;
    out[0] = u[0];
    out[1] = u[1];
}
static const double linear2_u0[2] = {1.0, 1.0};
static const double linear2_p[2] = {1.0, 2.0};
static const orc_pdf_dim linear2_pdf[2] = {
    {ORC_PDF_UNIFORM, 0, 1.0, 10.0, 0.0, 1.0},
    {ORC_PDF_TRUNCNORMAL, 0, 0.0, 6.0, 3.0, 1.0},
};

/* ---------- C3: Lotka-Volterra (docs/src/tutorials/gpu_bayesian.md:36-50, :83-86, :141-143, :168) */
static void lv_rhs(double *du, const double *u, const double *p, double t) {
    (void)t;
    double x = u[0], y = u[1];
    double al = p[0], be = p[1], ga = p[2], de = p[3];
    du[0] = (al - be * y) * x;
    du[1] = (de * x - ga) * y;
}
/* h(x,u,p) = u, x */
static void map_x_to_p4(const double *x, const double *u0n, const double *pn, double *u0, double *p) {
    (void)u0n; (void)pn; (void)u0;
    for (int i = 0; i < 4; i++) p[i] = x[i];
}
/* g(sol,p) = sol[1,end] */
static void obs_first(const double *u, const double *p, double t, double *out) {
    (void)p; (void)t;
    out[0] = u[0];
}
static const double lv_u0[2] = {1.0, 1.0};
static const double lv_p[4] = {1.5, 1.0, 3.0, 1.0};
static const orc_pdf_dim lv_pdf[4] = {
    {ORC_PDF_TRUNCNORMAL, 0, 1.0, 2.0, 1.5, 0.5},
    {ORC_PDF_TRUNCNORMAL, 0, 0.5, 1.5, 1.2, 0.5},
    {ORC_PDF_TRUNCNORMAL, 0, 2.0, 4.0, 3.0, 0.5},
    {ORC_PDF_TRUNCNORMAL, 0, 0.5, 1.5, 1.0, 0.5},
};

/* ---------- C4: Lorenz-63, 3 uncertain params + 3 uncertain u0 (synthetic) ---- */
static void lorenz_rhs(double *du, const double *u, const double *p, double t) {
    (void)t;
    du[0] = p[0] * (u[1] - u[0]);
    du[1] = u[0] * (p[1] - u[2]) - u[1];
    du[2] = u[0] * u[1] - p[2] * u[2];
}
static void map_lorenz(const double *x, const double *u0n, const double *pn, double *u0, double *p) {
    (void)u0n; (void)pn;
    p[0] = x[0]; p[1] = x[1]; p[2] = x[2];
    u0[0] = x[3]; u0[1] = x[4]; u0[2] = x[5];
}
static void obs_state3(const double *u, const double *p, double t, double *out) {
    (void)p; (void)t;
    out[0] = u[0]; out[1] = u[1]; out[2] = u[2];
}
static const double lorenz_u0[3] = {1.0, 0.0, 0.0};
static const double lorenz_p[3] = {10.0, 28.0, 8.0 / 3.0};
static const orc_pdf_dim lorenz_pdf[6] = {
    {ORC_PDF_UNIFORM, 0, 9.0, 11.0, 0.0, 1.0},
    {ORC_PDF_UNIFORM, 0, 25.2, 30.8, 0.0, 1.0},
    {ORC_PDF_UNIFORM, 0, 0.9 * (8.0 / 3.0), 1.1 * (8.0 / 3.0), 0.0, 1.0},
    {ORC_PDF_TRUNCNORMAL, 0, 0.6, 1.4, 1.0, 0.1},
    {ORC_PDF_TRUNCNORMAL, 0, -0.4, 0.4, 0.0, 0.1},
    {ORC_PDF_TRUNCNORMAL, 0, -0.4, 0.4, 0.0, 0.1},
};

/* ---------- test/solve.jl:137-173: du = -0.5 u, u0 ~ tN(1,0.1;0.5,1.5), tspan (0,2) -------- */
static void decay_rhs(double *du, const double *u, const double *p, double t) {
    (void)p; (void)t;
    du[0] = -0.5 * u[0];
}
static void map_x_to_u01(const double *x, const double *u0n, const double *pn, double *u0, double *p) {
    (void)u0n; (void)pn; (void)p;
    u0[0] = x[0];
}
static const double decay_u0[1] = {1.0};
static const orc_pdf_dim decay_pdf[1] = {{ORC_PDF_TRUNCNORMAL, 0, 0.5, 1.5, 1.0, 0.1}};

/* ---------- C5: damped oscillator with KKL process noise (noise construction
 * reference expectation.jl:304-314, problem_types.jl:81-85) --------------------------------- */
static void osc_drift(double *du, const double *u, const double *p, double t) {
    (void)t;
    double om = p[0], ze = p[1];
    du[0] = u[1];
    du[1] = -om * om * u[0] - 2.0 * ze * om * u[1];
}
static void osc_diff(double *g, const double *u, const double *p, double t) {
    (void)u; (void)t;
    g[0] = 0.0;
    g[1] = p[2];
}
/* cov(x,u,p) = x, p  ->  Z = x */
static void map_x_to_z8(const double *x, const double *u0n, const double *pn, double *z, double *p) {
    (void)u0n; (void)pn; (void)p;
    for (int k = 0; k < 8; k++) z[k] = x[k];
}
static const double osc_u0[2] = {1.0, 0.0};
static const double osc_p[3] = {6.283185307179586, 0.1, 1.0};
/* problem_types.jl:83: n x Truncated(Normal(), -4, 4) */
static const orc_pdf_dim kkl8_pdf[8] = {
    {ORC_PDF_TRUNCNORMAL, 0, -4.0, 4.0, 0.0, 1.0}, {ORC_PDF_TRUNCNORMAL, 0, -4.0, 4.0, 0.0, 1.0},
    {ORC_PDF_TRUNCNORMAL, 0, -4.0, 4.0, 0.0, 1.0}, {ORC_PDF_TRUNCNORMAL, 0, -4.0, 4.0, 0.0, 1.0},
    {ORC_PDF_TRUNCNORMAL, 0, -4.0, 4.0, 0.0, 1.0}, {ORC_PDF_TRUNCNORMAL, 0, -4.0, 4.0, 0.0, 1.0},
    {ORC_PDF_TRUNCNORMAL, 0, -4.0, 4.0, 0.0, 1.0}, {ORC_PDF_TRUNCNORMAL, 0, -4.0, 4.0, 0.0, 1.0},
};

/* ---------- test/processnoise.jl:9-14: GBM f = g = u, u0 = 1:4, scalar noise ---------------- */
static void gbm_f(double *du, const double *u, const double *p, double t) {
    (void)p; (void)t;
    for (int i = 0; i < 4; i++) du[i] = u[i];
}
static void obs_state4(const double *u, const double *p, double t, double *out) {
    (void)p; (void)t;
    for (int i = 0; i < 4; i++) out[i] = u[i];
}
static const double gbm_u0[4] = {1.0, 2.0, 3.0, 4.0};
static const double gbm_p[1] = {0.0};

int orc_builtin_problem(const char *name, orc_problem *pb) {
    memset(pb, 0, sizeof(*pb));
    pb->abstol = 1e-6; /* OrdinaryDiffEq defaults */
    pb->reltol = 1e-3;
    pb->maxiters = 100000;
    pb->solver = ORC_SOLVER_TSIT5;
    if (!strcmp(name, "linear2")) {
        pb->nstate = 2; pb->nparam = 2; pb->nx = 2; pb->nout = 2;
        pb->rhs = linear2_rhs; pb->hmap = map_x_to_u0; pb->obs = obs_state2;
        pb->u0_nom = linear2_u0; pb->p_nom = linear2_p; pb->t0 = 0.0; pb->t1 = 3.0;
        pb->pdf = linear2_pdf;
    } else if (!strcmp(name, "lotka_volterra")) {
        pb->nstate = 2; pb->nparam = 4; pb->nx = 4; pb->nout = 1;
        pb->rhs = lv_rhs; pb->hmap = map_x_to_p4; pb->obs = obs_first;
        pb->u0_nom = lv_u0; pb->p_nom = lv_p; pb->t0 = 0.0; pb->t1 = 10.0;
        pb->pdf = lv_pdf;
    } else if (!strcmp(name, "lorenz63")) {
        pb->nstate = 3; pb->nparam = 3; pb->nx = 6; pb->nout = 3;
        pb->rhs = lorenz_rhs; pb->hmap = map_lorenz; pb->obs = obs_state3;
        pb->u0_nom = lorenz_u0; pb->p_nom = lorenz_p; pb->t0 = 0.0; pb->t1 = 1.0;
        pb->pdf = lorenz_pdf;
    } else if (!strcmp(name, "decay1")) {
        pb->nstate = 1; pb->nparam = 0; pb->nx = 1; pb->nout = 1;
        pb->rhs = decay_rhs; pb->hmap = map_x_to_u01; pb->obs = obs_first;
        pb->u0_nom = decay_u0; pb->p_nom = gbm_p; pb->t0 = 0.0; pb->t1 = 2.0;
        pb->pdf = decay_pdf;
    } else if (!strcmp(name, "oscillator_noise")) {
        pb->nstate = 2; pb->nparam = 3; pb->nx = 8; pb->nout = 2;
        pb->solver = ORC_SOLVER_EM; pb->n_kkl = 8; pb->dt_fixed = 1.0 / 1024.0;
        pb->rhs = osc_drift; pb->diffusion = osc_diff; pb->hmap = map_x_to_z8; pb->obs = obs_state2;
        pb->u0_nom = osc_u0; pb->p_nom = osc_p; pb->t0 = 0.0; pb->t1 = 1.0;
        pb->pdf = kkl8_pdf;
    } else if (!strcmp(name, "gbm4")) {
        pb->nstate = 4; pb->nparam = 0; pb->nx = 8; pb->nout = 4;
        pb->solver = ORC_SOLVER_EM; pb->n_kkl = 8; pb->dt_fixed = 1.0 / 1024.0;
        pb->rhs = gbm_f; pb->diffusion = gbm_f; pb->hmap = map_x_to_z8; pb->obs = obs_state4;
        pb->u0_nom = gbm_u0; pb->p_nom = gbm_p; pb->t0 = 0.0; pb->t1 = 1.0;
        pb->pdf = kkl8_pdf;
    } else {
        return -1;
    }
    return 0;
}
