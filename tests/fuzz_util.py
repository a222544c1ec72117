"""Seeded random models for differential testing of the CUDA path against the oracle.

A model is a small dissipative ODE system (skew-symmetric coupling minus a positive diagonal, a bounded
nonlinearity, optional parameter-scaled damping and explicit time dependence) written once in the shared C / CUDA
dialect, so the kernel (through NVRTC) and the oracle (through gcc, ``oracle.OracleProblem.from_source``) compile the
same text; ``h`` scatters a random subset of x into u0 and p (the three forms of SURVEY.md 8a2), ``g`` picks random
end-state components plus one product -- or components at random save times -- and every x dimension gets a random
factor of the product density (uniform / truncated normal / normal)."""
import numpy as np

import scimlexpectations_jl_b200 as sme
from scimlexpectations_jl_b200._capi import PDF_NORMAL, PDF_TRUNCNORMAL, PDF_UNIFORM
from scipy.special import ndtr, ndtri


def _lit(v):
    return f"B200E_REAL({float(v)!r})"


def random_model(seed: int, nmax: int = 5):
    """Returns (spec, draw) with draw(n, seed) -> (n, nx) samples of the model's own density.  ``nmax`` > 5 draws
    the large systems (6..nmax states) whose stages do not fit the register file."""
    rng = np.random.default_rng([20261020, seed] + ([nmax] if nmax != 5 else []))
    n = int(rng.integers(1, 6)) if nmax <= 5 else int(rng.integers(6, nmax + 1))
    m = int(rng.integers(0, 4))
    S = np.triu(rng.uniform(-1.5, 1.5, (n, n)), 1)
    A = S - S.T - np.diag(rng.uniform(0.2, 1.5, n))
    lines = []
    for i in range(n):
        terms = [f"{_lit(A[i, j])} * u[{j}]" for j in range(n) if A[i, j] != 0.0]
        if rng.random() < 0.6:
            terms.append(f"{_lit(rng.uniform(-0.8, 0.8))} * sin(u[{(i + 1) % n}])")
        if m and rng.random() < 0.7:
            terms.append(f"- p[{i % m}] * u[{i}]")
        if rng.random() < 0.3:
            terms.append(f"{_lit(rng.uniform(-0.3, 0.3))} * cos(t)")
        lines.append(f"    du[{i}] = " + " + ".join(terms).replace("+ -", "- ") + ";")
    rhs = "B200E_FN void rhs(real* du, const real* u, const real* p, real t) {\n" + "\n".join(lines) + "\n}\n"

    # h: which components of u0 / p come from x (at least one)
    u_idx = [i for i in range(n) if rng.random() < 0.6]
    p_idx = [i for i in range(m) if rng.random() < 0.6]
    if not u_idx and not p_idx:
        u_idx = [0]
    nx = len(u_idx) + len(p_idx)
    hmap = sme.models.hmap_scatter(u0_from_x=[(i, j) for j, i in enumerate(u_idx)],
                                   p_from_x=[(i, len(u_idx) + j) for j, i in enumerate(p_idx)])
    pdf = []
    for j in range(nx):
        is_p = j >= len(u_idx)
        lo, hi = (0.5, 1.5) if is_p else sorted(rng.uniform(-2.0, 2.0, 2) + np.array([-0.2, 0.2]))
        kind = int(rng.choice([PDF_UNIFORM, PDF_TRUNCNORMAL, PDF_TRUNCNORMAL] + ([] if is_p else [PDF_NORMAL])))
        mu = float(rng.uniform(lo, hi))
        sg = float(rng.uniform(0.2, 0.8) * (hi - lo))
        if kind == PDF_NORMAL:
            pdf.append((PDF_NORMAL, -np.inf, np.inf, mu, 0.3 * sg))
        else:
            pdf.append((kind, float(lo), float(hi), mu, sg))

    T = float(rng.uniform(0.5, 4.0))
    abstol, reltol = [(1e-6, 1e-3), (1e-8, 1e-6), (1e-4, 1e-2)][int(rng.integers(0, 3))]
    comps = sorted(rng.choice(n, size=int(rng.integers(1, n + 1)), replace=False).tolist())
    save_times = ()
    if rng.random() < 0.3:
        save_times = tuple(sorted(rng.uniform(0.0, T, int(rng.integers(1, 5))).tolist()))
        if rng.random() < 0.5:
            save_times = save_times + (T,)
        obs = sme.models.observable_at_components(comps)
        nout = len(save_times) * len(comps)
    else:
        body = "\n".join(f"    out[{k}] = u[{i}];" for k, i in enumerate(comps))
        body += f"\n    out[{len(comps)}] = u[{comps[0]}] * u[{comps[-1]}]" + (" + p[0];" if m else ";")
        obs = ("B200E_FN void observable(const real* u, const real* p, real t_end, real* out) {\n" + body + "\n}\n")
        nout = len(comps) + 1
    schedule = int(rng.integers(0, 3))
    if save_times and schedule == 2:
        schedule = 1
    spec = sme.ModelSpec(source=rhs + hmap + obs, nstate=n, nparam=m, nx=nx, nout=nout,
                         u0=rng.uniform(-1.0, 1.0, n).tolist(), p=rng.uniform(0.5, 1.5, m).tolist(), tspan=(0.0, T),
                         abstol=abstol, reltol=reltol, pdf=pdf, schedule=schedule, save_times=save_times)

    def draw(npts: int, sseed: int = 0):
        g = np.random.default_rng([seed, sseed])
        u = g.random((npts, nx))
        x = np.empty_like(u)
        for j, (kind, lo, hi, mu, sg) in enumerate(pdf):
            if kind == PDF_UNIFORM:
                x[:, j] = lo + (hi - lo) * u[:, j]
            elif kind == PDF_NORMAL:
                x[:, j] = mu + sg * ndtri(u[:, j])
            else:
                a, b = ndtr((lo - mu) / sg), ndtr((hi - mu) / sg)
                x[:, j] = np.clip(mu + sg * ndtri(a + (b - a) * u[:, j]), lo, hi)
        return x

    return spec, draw


def with_logpoly_factors(seed: int, spec, draw):
    """The same model with the density factors of its positive-support dimensions (those scattered into p, support
    [0.5, 1.5]) replaced by log-polynomial ones (B200E_PDF_LOGPOLY): truncated Gamma / LogNormal / Exponential, or a
    Beta on (0.5, 1) -- drawn from a separate stream so that random_model(seed) itself is unchanged."""
    rng = np.random.default_rng([20261022, seed])
    pdf, facs = list(spec.pdf), {}
    for j, d in enumerate(spec.pdf):
        if (d[1], d[2]) != (0.5, 1.5):
            continue
        kind = int(rng.integers(0, 4))
        if kind == 0:
            f = sme.truncated(sme.Gamma(float(rng.uniform(0.6, 9.0)), float(rng.uniform(0.1, 1.0))), 0.5, 1.5)
        elif kind == 1:
            f = sme.truncated(sme.LogNormal(float(rng.uniform(-0.5, 0.5)), float(rng.uniform(0.2, 0.8))), 0.5, 1.5)
        elif kind == 2:
            f = sme.truncated(sme.Exponential(float(rng.uniform(0.3, 3.0))), 0.5, 1.5)
        else:
            f = sme.truncated(sme.Beta(float(rng.uniform(0.7, 4.0)), float(rng.uniform(0.7, 4.0))), 0.5, 1.0)
        pdf[j], facs[j] = f.table(), f
    if not facs:   # no parameter dimension: put a Gamma-shaped factor on a shifted copy of nothing -- leave the model as it is
        return spec, draw

    def draw2(npts: int, sseed: int = 0):
        x = draw(npts, sseed)
        u = np.random.default_rng([seed, sseed, 5]).random((npts, len(facs)))
        for c, (j, f) in enumerate(facs.items()):
            x[:, j] = f.icdf(u[:, c])
        return x

    return spec.replace(pdf=pdf), draw2


def random_noise_model(seed: int):
    """A process-noise model (src/expectation.jl:209-249 / :295-325): 1-4 states, random drift and diffusion
    (zero / constant / linear / bounded-nonlinear per component), 1-13 KKL modes (odd and even counts), fixed-step
    Euler-Maruyama or adaptive LambaEM; x = Z with the +-4 sigma truncated normals of problem_types.jl:83."""
    rng = np.random.default_rng([20261021, seed])
    n = int(rng.integers(1, 5))
    nk = int(rng.choice([1, 2, 3, 5, 8, 13]))
    S = np.triu(rng.uniform(-2.0, 2.0, (n, n)), 1)
    A = S - S.T - np.diag(rng.uniform(0.1, 1.0, n))
    fl, gl = [], []
    for i in range(n):
        terms = [f"{_lit(A[i, j])} * u[{j}]" for j in range(n) if A[i, j] != 0.0]
        if rng.random() < 0.4:
            terms.append(f"{_lit(rng.uniform(-0.5, 0.5))} * sin(u[{(i + 1) % n}] + t)")
        fl.append(f"    du[{i}] = " + " + ".join(terms).replace("+ -", "- ") + ";")
        kind = int(rng.integers(0, 4))
        c = _lit(rng.uniform(0.2, 1.0))
        gl.append(f"    g[{i}] = " + ["B200E_REAL(0.0)", c, f"{c} * u[{i}]", f"{c} * (B200E_REAL(1.0) + B200E_REAL(0.25) * cos(u[{i}]))"][kind] + ";")
    src = ("B200E_FN void rhs(real* du, const real* u, const real* p, real t) {\n" + "\n".join(fl) + "\n}\n"
           "B200E_FN void diffusion(real* g, const real* u, const real* p, real t) {\n" + "\n".join(gl) + "\n}\n")
    src += sme.models.hmap_scatter(z_from_x=[(k, k) for k in range(nk)])
    comps = sorted(rng.choice(n, size=int(rng.integers(1, n + 1)), replace=False).tolist())
    src += sme.models.observable_components(comps)
    adaptive = bool(rng.integers(0, 2))
    tol = float(rng.choice([1e-2, 1e-3]))
    pdf = [(PDF_TRUNCNORMAL, -4.0, 4.0, 0.0, 1.0)] * nk
    spec = sme.ModelSpec(source=src, nstate=n, nparam=0, nx=nk, nout=len(comps), u0=rng.uniform(0.5, 1.5, n).tolist(), p=[],
                         tspan=(0.0, float(rng.uniform(0.5, 1.5))), solver=2 if adaptive else 1, n_kkl=nk,
                         abstol=tol, reltol=tol, dt_fixed=0.0 if adaptive else 2.0 ** -int(rng.integers(5, 10)),
                         pdf=pdf, schedule=int(rng.integers(0, 3)))

    def draw(npts: int, sseed: int = 0):
        g = np.random.default_rng([seed, sseed, 1])
        a, b = ndtr(-4.0), ndtr(4.0)
        return np.clip(ndtri(a + (b - a) * g.random((npts, nk))), -4.0, 4.0)

    return spec, draw


def oracle_of(spec):
    import oracle
    opts = dict(dt_fixed=spec.dt_fixed) if spec.solver == 1 else dict(abstol=spec.abstol, reltol=spec.reltol)
    return oracle.OracleProblem.from_source(spec.source, nstate=spec.nstate, nparam=spec.nparam, nx=spec.nx, nout=spec.nout,
                                            u0=spec.u0, p=spec.p, tspan=spec.tspan, pdf=spec.pdf, save_times=spec.save_times,
                                            solver=spec.solver, n_kkl=spec.n_kkl, **opts)


def compare(h, spec, orc, x, layout: int, weight: bool):
    """Kernel (open handle h) against oracle on the points x.  Returns a dict of what was observed; raises
    AssertionError on a parity violation."""
    n = len(x)
    xin = x if layout == 0 else np.ascontiguousarray(x.T)
    ref, cref, sref = orc.eval_nodes(x, weight_by_pdf=weight, want_steps=True)
    dx, steps = h.eval_nodes(xin, layout=layout, weight_by_pdf=weight, want_steps=True)
    if layout == 1:
        dx = dx.T
    c = h.last_counters()
    flipped = steps != sref
    slack = max(1, int(1e-4 * n))
    assert c["nfail"] == cref["nfail"] and flipped.sum() <= slack and np.all(np.abs(steps - sref)[flipped] <= 1), (c, cref, int(flipped.sum()))
    if not flipped.any():
        assert c == cref, (c, cref)
    scale = np.abs(ref).max(axis=0, keepdims=True) + 1e-300
    err = (np.abs(dx - ref) / scale)[~flipped]
    assert np.median(err) <= 1e-10 and err.max() <= 2e-6, (float(np.median(err)), float(err.max()))
    so, s2o, co = orc.reduce_samples(x, weight_by_pdf=weight)
    s, s2, cr = h.reduce_samples(xin, layout=layout, weight_by_pdf=weight)
    assert cr == c
    if not flipped.any():
        bound = 1e-8 * np.sqrt(n * s2o) + 1e-300          # |sum g| <= sqrt(n sum g^2): 1e-8 of the values' scale
        assert np.all(np.abs(s - so) <= bound) and np.all(np.abs(s2 - s2o) <= 1e-8 * s2o + 1e-300), (s, so)
    return {"n": n, "flipped": int(flipped.sum()), "median": float(np.median(err)), "max": float(err.max()),
            "steps_mean": float(steps.mean())}
