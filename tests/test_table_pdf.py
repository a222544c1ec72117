"""Tabulated density factors: ``pdf(kde, x)`` of KernelDensity estimates inside
``GenericDistribution(pdf_func, rand_func, lb, ub)`` (reference docs/src/tutorials/gpu_bayesian.md:153-166).
Linear interpolation on the KDE's uniform grid, 0 outside; Koopman only."""
import numpy as np
import pytest
import scipy.stats as st

import oracle
import scimlexpectations_jl_b200 as sme
from tests.sampling_util import sample_points

PRIORS = [(1.5, 0.5, 1.0, 2.0), (1.2, 0.5, 0.5, 1.5), (3.0, 0.5, 2.0, 4.0), (1.0, 0.5, 0.5, 1.5)]  # gpu_bayesian.md:83-86


def _tabulated_lv(npoints=4097):
    """lotka_volterra with every factor of its density replaced by a table of the same truncated normal."""
    spec = sme.models.builtin("lotka_volterra")
    pdf = []
    for mu, sg, lo, hi in PRIORS:
        x = np.linspace(lo, hi, npoints)
        pdf.append((sme._capi.PDF_TABLE, lo, hi, 0.0, 1.0, st.truncnorm((lo - mu) / sg, (hi - mu) / sg, loc=mu, scale=sg).pdf(x)))
    return spec.replace(pdf=pdf), pdf


def test_oracle_table_pdf_interpolates_linearly():
    spec, pdf = _tabulated_lv(npoints=33)
    orc = oracle.OracleProblem.builtin("lotka_volterra")
    orc.set_pdf(pdf)
    rng = np.random.default_rng(0)
    for _ in range(200):
        x = np.array([rng.uniform(lo, hi) for _, _, lo, hi in PRIORS])
        ref = np.prod([np.interp(xj, np.linspace(d[1], d[2], 33), d[5]) for xj, d in zip(x, pdf)])
        assert np.isclose(np.exp(orc.logpdf(x)), ref, rtol=1e-13)
    assert orc.logpdf(np.array([2.5, 1.0, 3.0, 1.0])) == -np.inf       # outside the grid: density 0
    # grid ends belong to the support
    assert np.isfinite(orc.logpdf(np.array([1.0, 0.5, 2.0, 0.5]))) and np.isfinite(orc.logpdf(np.array([2.0, 1.5, 4.0, 1.5])))


def test_api_kde_matches_its_own_table_and_cannot_be_sampled():
    rng = np.random.default_rng(1)
    k = sme.UnivariateKDE.from_samples(rng.normal(1.5, 0.1, 4000), npoints=512)
    assert abs(np.trapezoid(k.density, k.x) - 1.0) < 1e-6
    assert k.pdf(k.x[100]) == k.density[100] and k.pdf(k.x[0] - 1.0) == 0.0
    assert abs(k.pdf(1.5) - st.norm(1.5, 0.1).pdf(1.5)) < 0.4       # a KDE of 4000 samples, not the exact density
    spec, _ = _tabulated_lv()
    with pytest.raises(sme.B200Error):
        sme.sample_host(spec, 1, 10)


@pytest.mark.gpu
def test_gpu_table_pdf_matches_oracle_and_the_analytic_density(gpu):
    spec, pdf = _tabulated_lv()
    x = sample_points("lotka_volterra", 20000, seed=31)
    orc = oracle.OracleProblem.builtin("lotka_volterra")
    orc.set_pdf(pdf)
    ref, cref = orc.eval_nodes(x, weight_by_pdf=True)
    with gpu.Handle(spec) as h:
        dx = h.eval_nodes(x, weight_by_pdf=True)
        c = h.last_counters()
        u_tab, r_tab, st_tab = h.koopman_solve([1.0, 0.5, 2.0, 0.5], [2.0, 1.5, 4.0, 1.5], reltol=1e-3, abstol=1e-3)
    assert c == cref
    scale = np.abs(ref).max()
    err = np.abs(dx - ref) / scale
    assert np.median(err) <= 1e-10 and err.max() <= 2e-6
    # against the same problem with the analytic truncated normals: only the interpolation error (h^2/8 f'') is left
    with gpu.Handle(gpu.models.builtin("lotka_volterra")) as h:
        dxa = h.eval_nodes(x, weight_by_pdf=True)
        u_an, r_an, st_an = h.koopman_solve([1.0, 0.5, 2.0, 0.5], [2.0, 1.5, 4.0, 1.5], reltol=1e-3, abstol=1e-3)
    assert np.max(np.abs(dx - dxa)) / scale < 1e-6
    assert abs(u_tab[0] - u_an[0]) < 1e-6 * abs(u_an[0]) and st_tab["nevals"] == st_an["nevals"]


@pytest.mark.gpu
def test_gpu_bayesian_tutorial_shape(gpu):
    """gpu_bayesian.md:153-170 through the mirrored API: four KDEs of (synthetic) posterior samples as the
    density, h(x,u,p) = u, x, g = sol[1, end], Koopman; checked against the oracle driven by the same tables."""
    m = gpu.models
    rng = np.random.default_rng(7)
    kdes = [gpu.UnivariateKDE.from_samples(rng.normal(mu, 0.05, 2000), npoints=1024) for mu in (1.5, 1.0, 3.0, 1.0)]
    prob = gpu.ODEProblem(m._LV_RHS, [1.0, 1.0], (0.0, 10.0), [1.5, 1.0, 3.0, 1.0])
    smap = gpu.SystemMap(prob, gpu.Tsit5(), gpu.EnsembleB200())
    gd = gpu.GenericDistribution(*kdes)
    ex = gpu.ExpectationProblem(smap, m.observable_components([0]), m.hmap_scatter(p_from_x=[(0, 0), (1, 1), (2, 2), (3, 3)]), gd, nout=1)
    sol = gpu.solve(ex, gpu.Koopman(), batch=10000, ireltol=1e-2, iabstol=1e-2)      # the tutorial's tolerances
    orc = oracle.OracleProblem.builtin("lotka_volterra")
    orc.set_pdf(gd.table())
    lb, ub = gpu.extrema(gd)
    ref = gpu.hcubature_v(lambda x: orc.eval_nodes(x, weight_by_pdf=True)[0], lb, ub, fdim=1, reltol=1e-2, abstol=1e-2, maxevals=1000000)
    assert abs(sol.u - ref.u[0]) < 1e-8 * abs(ref.u[0]) and sol.original.nevals == ref.nevals
    assert 0.5 < sol.u < 2.0
    ex.close()
