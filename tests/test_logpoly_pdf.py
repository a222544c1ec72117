"""Log-polynomial density factors (``B200E_PDF_LOGPOLY``, SURVEY.md 8 row a5 widened): the exponential-family
members of Distributions.jl a ``GenericDistribution`` is usually built from -- Exponential, Gamma, Beta, LogNormal
and their truncations -- as ``logpdf = c0 + c1 x + c2 x^2 + c3 log x + c4 log(1 - x) + c5 log(x)^2`` on the support.
Distributions.jl is a dependency of the reference, not part of it (Project.toml:30, compat "0.25.88");
its published log-densities are pinned here against ``scipy.stats`` as the independent implementation."""
import math

import numpy as np
import pytest
import scipy.stats as st

import oracle
import scimlexpectations_jl_b200 as sme
from tests.conftest import record_parity
from tests.sampling_util import sample_points

FAMILIES = [
    (lambda: sme.Exponential(0.7), st.expon(scale=0.7)),
    (lambda: sme.Gamma(2.5, 0.4), st.gamma(2.5, scale=0.4)),
    (lambda: sme.Gamma(0.5, 2.0), st.gamma(0.5, scale=2.0)),      # alpha < 1: unbounded at 0
    (lambda: sme.Beta(2.0, 3.5), st.beta(2.0, 3.5)),
    (lambda: sme.Beta(0.6, 0.8), st.beta(0.6, 0.8)),
    (lambda: sme.LogNormal(0.3, 0.45), st.lognorm(0.45, scale=math.exp(0.3))),
    (lambda: sme.LogNormal(0.2, math.sqrt(0.2)), st.lognorm(math.sqrt(0.2), scale=math.exp(0.2))),   # c3 == 0, c5 != 0
    (lambda: sme.Rayleigh(1.3), st.rayleigh(scale=1.3)),
    (lambda: sme.Chisq(5.0), st.chi2(5.0)),
    (lambda: sme.Chi(3.0), st.chi(3.0)),
    (lambda: sme.Erlang(3, 0.5), st.gamma(3.0, scale=0.5)),
    (lambda: sme.Pareto(2.5, 0.8), st.pareto(2.5, scale=0.8)),
]


def _one_dim(table):
    one = oracle.OracleProblem.builtin("decay1")
    one.set_pdf([table])
    return one


@pytest.mark.parametrize("k", range(len(FAMILIES)))
def test_oracle_and_mirror_logpdf_match_scipy(k):
    make, ref = FAMILIES[k]
    d = make()
    one = _one_dim(d.table())
    lo, hi = ref.ppf(1e-6), ref.ppf(1 - 1e-6)
    for x in np.random.default_rng(k).uniform(lo, hi, 300):
        want = ref.pdf(x)
        assert abs(one.pdf(np.array([x])) - want) <= 1e-12 * max(1.0, want)
        assert abs(math.exp(d.logpdf(x)) - want) <= 1e-12 * max(1.0, want)
    # outside the support: 0; at a support end-point the limit Distributions.jl returns (0, a finite value or Inf)
    assert one.pdf(np.array([d.minimum() - 0.1])) == 0.0 and d.logpdf(d.minimum() - 0.1) == -math.inf
    edge = one.pdf(np.array([d.minimum()]))
    assert edge == edge and edge == pytest.approx(ref.pdf(d.minimum()), rel=1e-12)
    assert math.exp(d.logpdf(d.minimum())) == pytest.approx(edge, rel=1e-12) if math.isfinite(edge) else d.logpdf(d.minimum()) == math.inf


def test_truncation_folds_the_normaliser_into_c0():
    g = sme.Gamma(3.0, 0.5)
    t = sme.truncated(g, 0.5, 2.5)
    assert isinstance(t, sme.TruncatedLogPoly) and (t.minimum(), t.maximum()) == (0.5, 2.5)
    mass = st.gamma(3.0, scale=0.5).cdf(2.5) - st.gamma(3.0, scale=0.5).cdf(0.5)
    one = _one_dim(t.table())
    for x in np.linspace(0.5, 2.5, 41):
        assert one.pdf(np.array([x])) == pytest.approx(st.gamma(3.0, scale=0.5).pdf(x) / mass, rel=1e-12)
    assert one.pdf(np.array([2.6])) == 0.0 and one.pdf(np.array([0.4])) == 0.0
    xs = np.linspace(0.5, 2.5, 20001)
    assert abs(np.trapezoid([math.exp(t.logpdf(x)) for x in xs], xs) - 1.0) < 1e-7
    u = np.random.default_rng(0).random(2000)
    s = t.icdf(u)
    assert s.min() >= 0.5 and s.max() <= 2.5 and np.allclose([t.cdf(x) for x in s[:50]], u[:50], atol=1e-12)
    # truncated(Normal) keeps its own device kind (the in-kernel sampler draws it)
    assert isinstance(sme.truncated(sme.Normal(0.0, 1.0), -1.0, 1.0), sme.Truncated)
    with pytest.raises(TypeError):
        sme.truncated(sme.Uniform(0.0, 1.0), 0.2, 0.4)
    with pytest.raises(ValueError):
        sme.truncated(sme.Beta(2.0, 2.0), 1.5, 2.0)


def _spec(pdf):
    return sme.models.builtin("lotka_volterra").replace(pdf=pdf)


def test_host_validation_of_logpoly_factors(sme):
    good = sme.truncated(sme.Gamma(2.0, 0.5), 1.0, 2.0).table()
    rest = [sme.Uniform(0.5, 1.5).table(), sme.Uniform(2.0, 4.0).table(), sme.Uniform(0.5, 1.5).table()]
    bad = [
        (sme._capi.PDF_LOGPOLY, -1.0, 2.0, 0.0, 1.0, good[5]),                            # log x on a support below 0
        (sme._capi.PDF_LOGPOLY, 0.0, 2.0, 0.0, 1.0, sme.Beta(2.0, 2.0).table()[5]),       # log(1 - x) above 1
        (sme._capi.PDF_LOGPOLY, 1.0, 2.0, 0.0, 1.0, np.array([0.0, 0.0, np.nan, 0, 0, 0])),
    ]
    for b in bad:
        with pytest.raises(sme.B200Error):
            sme.compile_check(_spec([b] + rest))
    with pytest.raises(ValueError):
        _spec([(sme._capi.PDF_LOGPOLY, 1.0, 2.0, 0.0, 1.0, np.zeros(5))] + rest).to_c()
    # no inverse CDF in the library: the samplers refuse, MonteCarlo falls back to host samples (api._solve_montecarlo)
    with pytest.raises(sme.B200Error):
        sme.sample_host(_spec([good] + rest), 1, 10)
    gd = sme.GenericDistribution(sme.truncated(sme.Gamma(2.0, 0.5), 1.0, 2.0), sme.Beta(2.0, 3.0))
    x = gd.rand_block(3, 0, 4000)
    assert x[:, 0].min() >= 1.0 and x[:, 0].max() <= 2.0 and 0.0 < x[:, 1].min() and x[:, 1].max() < 1.0
    assert abs(x[:, 1].mean() - 0.4) < 0.02
    assert np.array_equal(x, gd.rand_block(3, 0, 4000)) and not np.array_equal(x, gd.rand_block(3, 1, 4000))


def test_handle_cache_key_sees_the_coefficients():
    from scimlexpectations_jl_b200.api import _spec_key
    rest = [sme.Uniform(0.5, 1.5).table(), sme.Uniform(2.0, 4.0).table(), sme.Uniform(0.5, 1.5).table()]
    a = _spec([sme.truncated(sme.Gamma(2.0, 0.5), 1.0, 2.0).table()] + rest)
    b = _spec([sme.truncated(sme.Gamma(2.0, 0.6), 1.0, 2.0).table()] + rest)
    assert _spec_key(a) != _spec_key(b) and _spec_key(a) == _spec_key(_spec([sme.truncated(sme.Gamma(2.0, 0.5), 1.0, 2.0).table()] + rest))


# the Lotka-Volterra parameters under positive-support priors, truncated to the box the builtin integrates over
def _lv_priors():
    return [sme.truncated(sme.Gamma(9.0, 1.5 / 9.0), 1.0, 2.0), sme.truncated(sme.LogNormal(0.0, 0.4), 0.5, 1.5),
            sme.truncated(sme.Exponential(3.0), 2.0, 4.0), sme.Uniform(0.5, 1.5)]


@pytest.mark.gpu
def test_gpu_logpoly_pdf_matches_oracle(gpu):
    pri = _lv_priors()
    pdf = [d.table() for d in pri]
    x = sample_points("lotka_volterra", 20000, seed=41)
    orc = oracle.OracleProblem.builtin("lotka_volterra")
    orc.set_pdf(pdf)
    ref, cref = orc.eval_nodes(x, weight_by_pdf=True)
    lb, ub = [1.0, 0.5, 2.0, 0.5], [2.0, 1.5, 4.0, 1.5]
    with gpu.Handle(_spec(pdf)) as h:
        dx = h.eval_nodes(x, weight_by_pdf=True)
        c = h.last_counters()
        dx1 = h.eval_nodes(x, weight_by_pdf=False)
        s, _, cs = h.reduce_samples(x, weight_by_pdf=True)
        u, resid, stt = h.koopman_solve(lb, ub, reltol=1e-3, abstol=1e-3)
    assert c == cref and cs == cref
    scale = np.abs(ref).max()
    err = np.abs(dx - ref) / scale
    assert np.median(err) <= 1e-10 and err.max() <= 2e-6
    w_gpu = dx[:400, 0] / dx1[:400, 0]
    w_orc = np.array([orc.pdf(xi) for xi in x[:400]])
    werr = float(np.max(np.abs(w_gpu - w_orc) / np.abs(w_orc).max()))
    assert werr <= 1e-12
    so, _, _ = orc.reduce_samples(x, weight_by_pdf=True)
    assert np.max(np.abs(s - so)) <= 1e-9 * np.abs(so).max()
    want = gpu.hcubature_v(lambda y: orc.eval_nodes(y, weight_by_pdf=True)[0], lb, ub, fdim=orc.nout, reltol=1e-3, abstol=1e-3,
                           maxevals=1000000)
    assert stt["nevals"] == want.nevals and np.max(np.abs(u - want.u)) <= 1e-8 * np.abs(want.u).max()
    record_parity("logpoly_pdf", weight_rel=werr, nodes_rel_max=float(err.max()),
                  koopman_rel=float(np.max(np.abs(u - want.u)) / np.abs(want.u).max()))


@pytest.mark.gpu
def test_gpu_logpoly_priors_through_the_mirrored_api(gpu):
    """Koopman and MonteCarlo over a GenericDistribution of log-polynomial factors; MonteCarlo ships host samples
    because the device holds no inverse CDF for them."""
    m = gpu.models
    pri = _lv_priors()
    gd = gpu.GenericDistribution(*pri)
    prob = gpu.ODEProblem(m._LV_RHS, [1.0, 1.0], (0.0, 10.0), [1.5, 1.0, 3.0, 1.0])
    smap = gpu.SystemMap(prob, gpu.Tsit5(), gpu.EnsembleB200())
    ex = gpu.ExpectationProblem(smap, m.observable_components([0]), m.hmap_scatter(p_from_x=[(0, 0), (1, 1), (2, 2), (3, 3)]), gd, nout=1)
    sol = gpu.solve(ex, gpu.Koopman(), batch=10000, ireltol=1e-3, iabstol=1e-3)
    orc = oracle.OracleProblem.builtin("lotka_volterra")
    orc.set_pdf(gd.table())
    lb, ub = gpu.extrema(gd)
    ref = gpu.hcubature_v(lambda x: orc.eval_nodes(x, weight_by_pdf=True)[0], lb, ub, fdim=1, reltol=1e-3, abstol=1e-3, maxevals=1000000)
    assert abs(sol.u - ref.u[0]) < 1e-8 * abs(ref.u[0]) and sol.original.nevals == ref.nevals
    mc = gpu.solve(ex, gpu.MonteCarlo(40000, seed=5, block=16384))
    xs = np.concatenate([gd.rand_block(5, b, 16384) for b in range(3)])[:40000]
    so, _, co = orc.reduce_samples(xs, weight_by_pdf=False)
    assert mc.counters == co and abs(mc.u - so[0] / 40000) <= 1e-9 * abs(so[0] / 40000)
    assert abs(mc.u - sol.u) < 0.02 * abs(sol.u)      # two estimators of the same expectation
    ex.close()
