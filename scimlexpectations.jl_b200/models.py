"""CUDA model sources for the BASELINE.json configs (product side).

Each model is the text of the device functions the NVRTC translation unit needs (``rhs``, ``hmap``,
``observable`` and, for Euler-Maruyama, ``diffusion``) in the shared dialect documented in
``csrc/kernels/expect_prelude.cuh``.  The oracle carries its own, independently written, C versions
of the same problems (``oracle/models_builtin.c``), so a transcription slip on either side fails parity.

Reference sources of the problems:
  linear2          README.md:31-61, test/solve.jl:17-84            (C1, C2)
  lotka_volterra   docs/src/tutorials/gpu_bayesian.md:36-50, :83-86, :141-143, :168   (C3)
  lorenz63         synthetic
  decay1           test/solve.jl:137-173
  oscillator_noise synthetic, test/processnoise.jl-style (C5); noise: src/expectation.jl:304-314
  gbm4             test/processnoise.jl:9-26
"""
from __future__ import annotations

import math

from ._capi import (ModelSpec, PDF_TRUNCNORMAL, PDF_UNIFORM, SOLVER_EM, SOLVER_LAMBA_EM, SOLVER_TSIT5)

# ---- reusable pieces: the structured h / g forms the Julia shim recognises (SURVEY.md hard part 8) ----


def hmap_scatter(u0_from_x=(), p_from_x=(), z_from_x=()):
    """``h(x,u,p)`` that scatters components of x into u0 / p (or Z for process noise).

    ``u0_from_x = [(i_u0, j_x), ...]`` means ``u0[i_u0] = x[j_x]``; everything not listed keeps the
    nominal value (the kernel pre-fills u0 / p with ``prob.u0`` / ``prob.p``)."""
    lines = []
    for i, j in u0_from_x:
        lines.append(f"    u0[{int(i)}] = x[{int(j)}];")
    for i, j in z_from_x:
        lines.append(f"    u0[{int(i)}] = x[{int(j)}];")
    for i, j in p_from_x:
        lines.append(f"    p[{int(i)}] = x[{int(j)}];")
    body = "\n".join(lines)
    return ("B200E_FN void hmap(const real* x, const real* u0_nom, const real* p_nom, real* u0, real* p) {\n"
            f"{body}\n}}\n")


def observable_components(idx):
    """``g(sol,p) = sol[idx, end]`` (end-state components)."""
    body = "\n".join(f"    out[{k}] = u[{int(i)}];" for k, i in enumerate(idx))
    return ("B200E_FN void observable(const real* u, const real* p, real t_end, real* out) {\n"
            f"{body}\n}}\n")


def observable_at_components(idx):
    """Time-sampled observable: ``g(sol,p) = sol[idx, :]`` with ``saveat`` resp. ``[sol(t_k)[i] for ...]``
    (docs/src/tutorials/introduction.md:53, :192, :207-212): slot k holds the components ``idx`` of the
    state at the k-th save time (``ModelSpec.save_times``); ``nout = len(save_times) * len(idx)``."""
    body = "\n".join(f"    out[{j}] = u[{int(i)}];" for j, i in enumerate(idx))
    return ("B200E_FN void observable_at(int k, real t, const real* u, const real* p, real* out) {\n"
            f"{body}\n}}\n")


# ---- C1 / C2 ----------------------------------------------------------------------------------
# f(du,u,p,t) with A = [0 1; -p1 -p2] captured at construction from the nominal p = [1, 2]
# (README.md:38-39): the matrix does not follow the remade parameters.
_LINEAR2_RHS = """
B200E_FN void rhs(real* du, const real* u, const real* p, real t) {
    du[0] = u[1];
    du[1] = -u[0] - B200E_REAL(2.0) * u[1];
}
"""


def linear2(**kw) -> ModelSpec:
    src = _LINEAR2_RHS + hmap_scatter(u0_from_x=[(0, 0), (1, 1)]) + observable_components([0, 1])
    return ModelSpec(source=src, nstate=2, nparam=2, nx=2, nout=2, u0=[1.0, 1.0], p=[1.0, 2.0],
                     tspan=(0.0, 3.0),
                     pdf=[(PDF_UNIFORM, 1.0, 10.0, 0.0, 1.0), (PDF_TRUNCNORMAL, 0.0, 6.0, 3.0, 1.0)]).replace(**kw)


def linear2_saveat(**kw) -> ModelSpec:
    """The README problem observed at six times (``g(sol,p) = sol[:, :]`` with saveat): both components."""
    ts = (0.0, 0.5, 1.0, 1.7, 2.5, 3.0)
    src = _LINEAR2_RHS + hmap_scatter(u0_from_x=[(0, 0), (1, 1)]) + observable_at_components([0, 1])
    return ModelSpec(source=src, nstate=2, nparam=2, nx=2, nout=2 * len(ts), u0=[1.0, 1.0], p=[1.0, 2.0],
                     tspan=(0.0, 3.0), save_times=ts,
                     pdf=[(PDF_UNIFORM, 1.0, 10.0, 0.0, 1.0), (PDF_TRUNCNORMAL, 0.0, 6.0, 3.0, 1.0)]).replace(**kw)


# docs/src/tutorials/introduction.md:14-24, :207-212: u' = p u, u0 ~ Uniform(0, 10), p = -0.3, tspan (0, 10),
# saveat = 0:0.5:10, g(sol, p) = sol[1, :]; analytical E = exp(p t) * mean(u0)
_SCALAR_GROWTH_RHS = """
B200E_FN void rhs(real* du, const real* u, const real* p, real t) {
    du[0] = p[0] * u[0];
}
"""


def growth_saveat(**kw) -> ModelSpec:
    ts = tuple(0.5 * k for k in range(21))
    src = _SCALAR_GROWTH_RHS + hmap_scatter(u0_from_x=[(0, 0)]) + observable_at_components([0])
    return ModelSpec(source=src, nstate=1, nparam=1, nx=1, nout=len(ts), u0=[10.0], p=[-0.3], tspan=(0.0, 10.0),
                     save_times=ts, pdf=[(PDF_UNIFORM, 0.0, 10.0, 0.0, 1.0)]).replace(**kw)


# ---- C3 ---------------------------------------------------------------------------------------
_LV_RHS = """
B200E_FN void rhs(real* du, const real* u, const real* p, real t) {
    const real x = u[0], y = u[1];
    du[0] = (p[0] - p[1] * y) * x;
    du[1] = (p[3] * x - p[2]) * y;
}
"""


def lotka_volterra(**kw) -> ModelSpec:
    src = _LV_RHS + hmap_scatter(p_from_x=[(0, 0), (1, 1), (2, 2), (3, 3)]) + observable_components([0])
    return ModelSpec(source=src, nstate=2, nparam=4, nx=4, nout=1, u0=[1.0, 1.0], p=[1.5, 1.0, 3.0, 1.0],
                     tspan=(0.0, 10.0),
                     pdf=[(PDF_TRUNCNORMAL, 1.0, 2.0, 1.5, 0.5), (PDF_TRUNCNORMAL, 0.5, 1.5, 1.2, 0.5),
                          (PDF_TRUNCNORMAL, 2.0, 4.0, 3.0, 0.5), (PDF_TRUNCNORMAL, 0.5, 1.5, 1.0, 0.5)]).replace(**kw)


# ---- C4 ---------------------------------------------------------------------------------------
_LORENZ_RHS = """
B200E_FN void rhs(real* du, const real* u, const real* p, real t) {
    du[0] = p[0] * (u[1] - u[0]);
    du[1] = u[0] * (p[1] - u[2]) - u[1];
    du[2] = u[0] * u[1] - p[2] * u[2];
}
"""


def lorenz63(**kw) -> ModelSpec:
    src = (_LORENZ_RHS + hmap_scatter(p_from_x=[(0, 0), (1, 1), (2, 2)], u0_from_x=[(0, 3), (1, 4), (2, 5)])
           + observable_components([0, 1, 2]))
    b = 8.0 / 3.0
    return ModelSpec(source=src, nstate=3, nparam=3, nx=6, nout=3, u0=[1.0, 0.0, 0.0], p=[10.0, 28.0, b],
                     tspan=(0.0, 1.0),
                     pdf=[(PDF_UNIFORM, 9.0, 11.0, 0.0, 1.0), (PDF_UNIFORM, 25.2, 30.8, 0.0, 1.0),
                          (PDF_UNIFORM, 0.9 * b, 1.1 * b, 0.0, 1.0),
                          (PDF_TRUNCNORMAL, 0.6, 1.4, 1.0, 0.1), (PDF_TRUNCNORMAL, -0.4, 0.4, 0.0, 0.1),
                          (PDF_TRUNCNORMAL, -0.4, 0.4, 0.0, 0.1)]).replace(**kw)


# ---- test/solve.jl:137-173 ----------------------------------------------------------------------
_DECAY_RHS = """
B200E_FN void rhs(real* du, const real* u, const real* p, real t) {
    du[0] = B200E_REAL(-0.5) * u[0];
}
"""


def decay1(**kw) -> ModelSpec:
    src = _DECAY_RHS + hmap_scatter(u0_from_x=[(0, 0)]) + observable_components([0])
    return ModelSpec(source=src, nstate=1, nparam=0, nx=1, nout=1, u0=[1.0], p=[], tspan=(0.0, 2.0),
                     pdf=[(PDF_TRUNCNORMAL, 0.5, 1.5, 1.0, 0.1)]).replace(**kw)


# ---- C5 ---------------------------------------------------------------------------------------
_OSC = """
B200E_FN void rhs(real* du, const real* u, const real* p, real t) {
    du[0] = u[1];
    du[1] = -(p[0] * p[0]) * u[0] - B200E_REAL(2.0) * p[1] * p[0] * u[1];
}
B200E_FN void diffusion(real* g, const real* u, const real* p, real t) {
    g[0] = B200E_REAL(0.0);
    g[1] = p[2];
}
"""

_KKL8_PDF = [(PDF_TRUNCNORMAL, -4.0, 4.0, 0.0, 1.0)] * 8  # problem_types.jl:83


def oscillator_noise(**kw) -> ModelSpec:
    src = _OSC + hmap_scatter(z_from_x=[(k, k) for k in range(8)]) + observable_components([0, 1])
    return ModelSpec(source=src, nstate=2, nparam=3, nx=8, nout=2, u0=[1.0, 0.0], p=[2.0 * math.pi, 0.1, 1.0],
                     tspan=(0.0, 1.0), solver=SOLVER_EM, n_kkl=8, dt_fixed=1.0 / 1024.0,
                     pdf=list(_KKL8_PDF)).replace(**kw)


# ---- test/processnoise.jl:9-14 --------------------------------------------------------------------
_GBM = """
B200E_FN void rhs(real* du, const real* u, const real* p, real t) {
    for (int i = 0; i < 4; i++) du[i] = u[i];
}
B200E_FN void diffusion(real* g, const real* u, const real* p, real t) {
    for (int i = 0; i < 4; i++) g[i] = u[i];
}
"""


def gbm4(**kw) -> ModelSpec:
    src = _GBM + hmap_scatter(z_from_x=[(k, k) for k in range(8)]) + observable_components([0, 1, 2, 3])
    return ModelSpec(source=src, nstate=4, nparam=0, nx=8, nout=4, u0=[1.0, 2.0, 3.0, 4.0], p=[],
                     tspan=(0.0, 1.0), solver=SOLVER_EM, n_kkl=8, dt_fixed=1.0 / 1024.0,
                     pdf=list(_KKL8_PDF)).replace(**kw)


def gbm4_lamba(**kw) -> ModelSpec:
    """test/processnoise.jl:9-15 as the reference runs it: ProcessNoiseSystemMap(prob, 8, LambaEM(),
    abstol = 1e-3, reltol = 1e-3) -- the adaptive solver instead of north_star's fixed-step Euler-Maruyama."""
    return gbm4(solver=SOLVER_LAMBA_EM, abstol=1e-3, reltol=1e-3, dt_fixed=0.0).replace(**kw)


# ---- a larger state: radioactive-decay chain u_0 -> u_1 -> ... -> u_7 with uncertain rates (synthetic; exercises
# the register budget: 8 states x 7 stages do not fit 128 registers) ----
def decay_chain8(**kw) -> ModelSpec:
    n = 8
    lines = ["    du[0] = -p[0] * u[0];"] + [f"    du[{i}] = p[{i - 1}] * u[{i - 1}] - p[{i}] * u[{i}];" for i in range(1, n)]
    rhs = "B200E_FN void rhs(real* du, const real* u, const real* p, real t) {\n" + "\n".join(lines) + "\n}\n"
    src = rhs + hmap_scatter(p_from_x=[(i, i) for i in range(n)]) + observable_components(list(range(n)))
    return ModelSpec(source=src, nstate=n, nparam=n, nx=n, nout=n, u0=[1.0] + [0.0] * (n - 1), p=[1.0] * n,
                     tspan=(0.0, 2.0),
                     pdf=[(PDF_UNIFORM, 0.5 + 0.1 * i, 1.5 + 0.1 * i, 0.0, 1.0) for i in range(n)]).replace(**kw)


BUILTIN = {
    "linear2": linear2,
    "lotka_volterra": lotka_volterra,
    "lorenz63": lorenz63,
    "decay1": decay1,
    "oscillator_noise": oscillator_noise,
    "gbm4": gbm4,
    "gbm4_lamba": gbm4_lamba,
    "linear2_saveat": linear2_saveat,
    "growth_saveat": growth_saveat,
    "decay_chain8": decay_chain8,
}


def builtin(name: str, **kw) -> ModelSpec:
    return BUILTIN[name](**kw)


__all__ = ["BUILTIN", "builtin", "hmap_scatter", "observable_components", "SOLVER_TSIT5", "SOLVER_EM", "SOLVER_LAMBA_EM"] + list(BUILTIN)
