"""Tabulated density factors: ``pdf(kde, x)`` of KernelDensity estimates inside
``GenericDistribution(pdf_func, rand_func, lb, ub)`` (reference docs/src/tutorials/gpu_bayesian.md:153-166).
Evaluated as KernelDensity's ``pdf(kde, x)`` does: the quadratic B-spline through the grid values
(``BSpline(Quadratic(Line(OnGrid())))``), 0 outside; Koopman only."""
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


def _interpolations_system(f):
    """The prefiltering system of Interpolations.jl for Quadratic(Line(OnGrid())), written out densely: an
    independent route to the coefficients (numpy's LU instead of the oracle's banded elimination)."""
    n = len(f)
    M = np.zeros((n + 2, n + 2))
    M[0, :3] = [1, -2, 1]
    M[-1, -3:] = [1, -2, 1]
    for i in range(1, n + 1):
        M[i, i - 1:i + 2] = [0.125, 0.75, 0.125]
    return np.linalg.solve(M, np.r_[0.0, f, 0.0])


def _spline_eval(c, lo, hi, n, x):
    g = (x - lo) * (n - 1) / (hi - lo)
    i = min(int(np.floor(g + 0.5)), n - 1)
    d = g - i
    return 0.5 * (d - 0.5) ** 2 * c[i] + (0.75 - d * d) * c[i + 1] + 0.5 * (d + 0.5) ** 2 * c[i + 2]


def test_oracle_kde_pdf_is_the_quadratic_bspline_of_kerneldensity():
    """pdf(kde, x) = the C1 quadratic B-spline through the grid values with zero second derivative in the end cells
    (KernelDensity's InterpKDE default), 0 outside the grid."""
    rng = np.random.default_rng(3)
    for n in (2, 3, 4, 33, 512):
        f = rng.uniform(0.1, 2.0, n)
        c = oracle.kde_prefilter(f)
        assert np.allclose(c, _interpolations_system(f), rtol=1e-13, atol=1e-14)
    spec, pdf = _tabulated_lv(npoints=33)
    orc = oracle.OracleProblem.builtin("lotka_volterra")
    orc.set_pdf(pdf)
    coefs = [_interpolations_system(d[5]) for d in pdf]
    for _ in range(200):
        x = np.array([rng.uniform(lo, hi) for _, _, lo, hi in PRIORS])
        ref = np.prod([_spline_eval(c, d[1], d[2], 33, xj) for xj, d, c in zip(x, pdf, coefs)])
        assert np.isclose(orc.pdf(x), ref, rtol=1e-13) and np.isclose(np.exp(orc.logpdf(x)), ref, rtol=1e-13)
    assert orc.logpdf(np.array([2.5, 1.0, 3.0, 1.0])) == -np.inf and orc.pdf(np.array([2.5, 1.0, 3.0, 1.0])) == 0.0
    # grid ends belong to the support
    assert np.isfinite(orc.logpdf(np.array([1.0, 0.5, 2.0, 0.5]))) and np.isfinite(orc.logpdf(np.array([2.0, 1.5, 4.0, 1.5])))
    # properties of the spline in one dimension (a decay1-shaped problem: one tabulated factor)
    lo, hi, n = -1.0, 2.0, 17
    grid = np.linspace(lo, hi, n)
    f = np.exp(-grid ** 2) + 0.3 * np.sin(3 * grid) ** 2
    one = oracle.OracleProblem.builtin("decay1")
    one.set_pdf([(4, lo, hi, 0.0, 1.0, f)])
    val = lambda x: one.pdf(np.array([x]))
    h = (hi - lo) / (n - 1)
    assert np.allclose([val(g) for g in grid], f, rtol=1e-14)                 # interpolates the grid values
    e = 1e-6 * h
    for k in range(n - 1):                                                      # C1 across the piece boundaries (cell midpoints)
        m = grid[k] + 0.5 * h
        assert abs(val(m + e) - val(m - e)) < 1e-5 * h
        dl = (val(m) - val(m - 2 * e)) / (2 * e)
        dr = (val(m + 2 * e) - val(m)) / (2 * e)
        assert abs(dl - dr) < 1e-3 * np.abs(f).max() / h
    for x0 in (lo + 0.2 * h, hi - 0.2 * h):                                     # zero second derivative in the end cells
        d2 = (val(x0 + 0.1 * h) - 2 * val(x0) + val(x0 - 0.1 * h)) / (0.1 * h) ** 2
        assert abs(d2) < 1e-8 * np.abs(f).max() / h ** 2
    lin = oracle.OracleProblem.builtin("decay1")
    lin.set_pdf([(4, lo, hi, 0.0, 1.0, 2.0 + 0.5 * grid)])                      # a straight line is reproduced exactly
    for x0 in rng.uniform(lo, hi, 50):
        assert abs(lin.pdf(np.array([x0])) - (2.0 + 0.5 * x0)) < 1e-13
    # against a smooth density the spline is third-order accurate: far closer than linear interpolation
    fine = np.linspace(lo + h, hi - h, 400)
    smooth = lambda x: np.exp(-x ** 2)
    sm = oracle.OracleProblem.builtin("decay1")
    sm.set_pdf([(4, lo, hi, 0.0, 1.0, smooth(grid))])
    err_spline = max(abs(sm.pdf(np.array([x0])) - smooth(x0)) for x0 in fine)
    err_linear = max(abs(np.interp(x0, grid, smooth(grid)) - smooth(x0)) for x0 in fine)
    assert err_spline < 0.25 * err_linear


def test_api_kde_matches_its_own_table_and_cannot_be_sampled():
    rng = np.random.default_rng(1)
    k = sme.UnivariateKDE.from_samples(rng.normal(1.5, 0.1, 4000), npoints=512)
    assert abs(np.trapezoid(k.density, k.x) - 1.0) < 1e-6
    assert abs(k.pdf(k.x[100]) - k.density[100]) < 1e-13 and k.pdf(k.x[0] - 1.0) == 0.0
    # the mirror's pdf(kde, x) is the oracle's, point for point
    one = oracle.OracleProblem.builtin("decay1")
    one.set_pdf([k.table()])
    for x0 in rng.uniform(k.x[0], k.x[-1], 100):
        assert abs(k.pdf(x0) - one.pdf(np.array([x0]))) <= 1e-13 * k.density.max()
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
        dx1 = h.eval_nodes(x, weight_by_pdf=False)
    assert c == cref
    scale = np.abs(ref).max()
    err = np.abs(dx - ref) / scale
    assert np.median(err) <= 1e-10 and err.max() <= 2e-6
    # the density alone (weighted / unweighted output of the same trajectories) against the oracle's pdf(kde, x):
    # host prefilter + three device taps reproduce the spline to rounding
    w_gpu = dx[:400, 0] / dx1[:400, 0]
    w_orc = np.array([orc.pdf(xi) for xi in x[:400]])
    assert np.max(np.abs(w_gpu - w_orc) / np.abs(w_orc).max()) <= 1e-12
    # against the same problem with the analytic truncated normals: only the interpolation error is left
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
