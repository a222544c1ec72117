"""Seeded sample sets for the builtin configs, shared by the CPU and GPU tests and by bench.py.

Blocks are keyed by (seed, block) so any block of a 1e9-sample stream can be regenerated alone
(SURVEY.md 8d C2): ``numpy.random.Generator(PCG64(SeedSequence([seed, block])))``, uniforms mapped
through the inverse CDF of each factor of the product density."""
import numpy as np
from scipy.special import ndtr, ndtri

BLOCK = 1 << 20

# (kind, lo, hi, mu, sigma) per dimension -- kept in step with the pdf tables of the models
_B = 8.0 / 3.0
DISTS = {
    "linear2": [("u", 1.0, 10.0), ("tn", 0.0, 6.0, 3.0, 1.0)],
    "lotka_volterra": [("tn", 1.0, 2.0, 1.5, 0.5), ("tn", 0.5, 1.5, 1.2, 0.5), ("tn", 2.0, 4.0, 3.0, 0.5),
                       ("tn", 0.5, 1.5, 1.0, 0.5)],
    "lorenz63": [("u", 9.0, 11.0), ("u", 25.2, 30.8), ("u", 0.9 * _B, 1.1 * _B), ("tn", 0.6, 1.4, 1.0, 0.1),
                 ("tn", -0.4, 0.4, 0.0, 0.1), ("tn", -0.4, 0.4, 0.0, 0.1)],
    "decay1": [("tn", 0.5, 1.5, 1.0, 0.1)],
    "oscillator_noise": [("tn", -4.0, 4.0, 0.0, 1.0)] * 8,
    "gbm4": [("tn", -4.0, 4.0, 0.0, 1.0)] * 8,
    "gbm4_lamba": [("tn", -4.0, 4.0, 0.0, 1.0)] * 8,
    "linear2_saveat": [("u", 1.0, 10.0), ("tn", 0.0, 6.0, 3.0, 1.0)],
    "growth_saveat": [("u", 0.0, 10.0)],
    "decay_chain8": [("u", 0.5 + 0.1 * i, 1.5 + 0.1 * i) for i in range(8)],
}


def _icdf(d, u):
    if d[0] == "u":
        return d[1] + (d[2] - d[1]) * u
    _, lo, hi, mu, sg = d
    a, b = ndtr((lo - mu) / sg), ndtr((hi - mu) / sg)
    return np.clip(mu + sg * ndtri(a + (b - a) * u), lo, hi)


def sample_block(name, seed, block, n=BLOCK):
    """Block ``block`` of the stream: (n, nx) float64, point-major."""
    dists = DISTS[name]
    rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence([int(seed), int(block)])))
    u = rng.random((n, len(dists)))
    out = np.empty_like(u)
    for j, d in enumerate(dists):
        out[:, j] = _icdf(d, u[:, j])
    return out


def sample_points(name, n, seed=0):
    """First n samples of the stream (whole blocks + a truncated last block)."""
    blocks = []
    got, b = 0, 0
    while got < n:
        m = min(BLOCK, n - got)
        blocks.append(sample_block(name, seed, b, BLOCK)[:m] if m < BLOCK else sample_block(name, seed, b))
        got += m
        b += 1
    return np.ascontiguousarray(np.concatenate(blocks, axis=0)) if len(blocks) > 1 else np.ascontiguousarray(blocks[0])


def oracle_for(name, **kw):
    """The oracle twin of a builtin product model.  The six BASELINE problems have independent C versions
    in oracle/models_builtin.c; the others (time-sampled observables, decay_chain8) are compiled from the model's own dialect source
    (the solver and the interpolant are still the oracle's own code)."""
    import oracle
    import scimlexpectations_jl_b200 as sme
    spec = sme.models.builtin(name)
    if not len(spec.save_times) and name in ("linear2", "lotka_volterra", "lorenz63", "decay1", "oscillator_noise", "gbm4"):
        return oracle.OracleProblem.builtin(name, **kw)
    opts = dict(abstol=spec.abstol, reltol=spec.reltol) if spec.solver != sme.models.SOLVER_EM else dict(dt_fixed=spec.dt_fixed)
    opts.update(kw)
    return oracle.OracleProblem.from_source(spec.source, nstate=spec.nstate, nparam=spec.nparam, nx=spec.nx, nout=spec.nout,
                                            u0=spec.u0, p=spec.p, tspan=spec.tspan, pdf=spec.pdf, solver=spec.solver,
                                            n_kkl=spec.n_kkl, save_times=spec.save_times, **opts)
