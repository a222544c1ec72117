"""Adaptive Euler-Maruyama (`LambaEM`) on the process-noise path: SURVEY.md 8(f) rank 4, second half.

The reference's own process-noise test (test/processnoise.jl:9-26) runs
``ProcessNoiseSystemMap(prob, 8, LambaEM(), abstol = 1e-3, reltol = 1e-3)`` on du = u dt + u dW, u0 = 1:4, and
asserts only that two expectations computed with that solver agree at rtol = 1e-2.  StochasticDiffEq is not
vendored, so the step sequence of the restatement is PARITY UNPINNED (oracle/expect_oracle.c, lamba_solve);
what is pinned here:
  CPU  the restated scheme converges (first order) to the closed form u0 exp(t + W(t)) the smooth KKL path
       admits, its controller obeys the stated clamps, and the model validates / compiles for sm_100a;
  GPU  the kernel reproduces the oracle on identical inputs: integer counters and per-point steps exactly,
       per-node values and sums within 1e-8 (same bars as the ODE path), under every schedule and with the
       in-kernel sampler; and MonteCarlo reproduces the closed-form expectation within the scheme's bias.
"""
import math

import numpy as np
import pytest

import oracle
from tests.sampling_util import oracle_for, sample_points

import os

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_vectors.npz"))
Z0 = np.array([[0.3, -1.2, 0.7, 0.1, -0.4, 2.0, -0.9, 0.05]])


def _w8(z, t):
    return math.sqrt(2.0) * sum(z[k] * math.sin((k + 0.5) * math.pi * t) / ((k + 0.5) * math.pi) for k in range(8))


def test_oracle_lamba_converges_to_the_closed_form():
    exact = np.array([1.0, 2.0, 3.0, 4.0]) * math.exp(1.0 + _w8(Z0[0], 1.0))
    errs, steps = [], []
    for tol in (1e-3, 1e-4, 1e-5, 1e-6):
        orc = oracle.OracleProblem.builtin("gbm4", solver=oracle.SOLVER_LAMBA_EM, abstol=tol, reltol=tol)
        dx, c = orc.eval_nodes(Z0, weight_by_pdf=False)
        assert c["nfail"] == 0 and c["nf"] == 2 + 2 * (c["naccept"] + c["nreject"])
        errs.append(float(np.max(np.abs(dx[0] - exact) / exact)))
        steps.append(c["naccept"])
    # error ~ mean step ~ sqrt(tol): each factor 10 in tol buys about sqrt(10) in error and costs as much in steps
    assert errs[0] < 0.06 and errs[-1] < 2e-3
    assert all(2.0 < a / b < 5.0 for a, b in zip(errs, errs[1:])), errs
    assert all(1.3 < b / a < 4.0 for a, b in zip(steps, steps[1:])), steps


def test_oracle_lamba_steps_obey_the_controller_clamps():
    """qmax = 1.125: from the tiny sde_determine_initdt step the accepted steps grow by at most 12.5 % each, so a
    solve from dt0 ~ 1e-7 needs at least log(1/dt0)/log(1.125) steps; maxiters aborts are counted, not thrown."""
    orc = oracle.OracleProblem.builtin("gbm4", solver=oracle.SOLVER_LAMBA_EM, abstol=1e-3, reltol=1e-3)
    _, c, steps = orc.eval_nodes(Z0, weight_by_pdf=False, want_steps=True)
    assert 60 <= steps[0] <= 1000 and c["nreject"] <= 0.05 * c["naccept"]
    orc.configure(maxiters=50)
    _, c = orc.eval_nodes(Z0, weight_by_pdf=False)
    assert c["nfail"] == 1 and c["naccept"] + c["nreject"] == 50


def test_oracle_lamba_reproduces_the_committed_fixture():
    """tests/golden/oracle_vectors.npz, entries gbm4_lamba_* (tests/golden/make_golden.py): the oracle built on
    this machine, in both of its forms (independent C model / compiled from the model's dialect source)."""
    x = GOLD["gbm4_lamba_x"]
    for orc in (oracle.OracleProblem.builtin("gbm4", solver=oracle.SOLVER_LAMBA_EM, abstol=1e-3, reltol=1e-3),
                oracle_for("gbm4_lamba")):
        dx, c, steps = orc.eval_nodes(x, weight_by_pdf=True, want_steps=True)
        assert np.array_equal(steps, GOLD["gbm4_lamba_steps"])
        assert [c["nf"], c["naccept"], c["nreject"], c["nfail"]] == GOLD["gbm4_lamba_counters"].tolist()
        assert np.allclose(dx, GOLD["gbm4_lamba_dx"], rtol=1e-12, atol=0)
        s, s2, _ = orc.reduce_samples(x)
        assert np.allclose(s, GOLD["gbm4_lamba_sum"], rtol=1e-13) and np.allclose(s2, GOLD["gbm4_lamba_sumsq"], rtol=1e-13)


def test_host_cubature_twin_in_eight_dimensions(sme):
    """The Koopman leg of the reference's process-noise test integrates over the 8 KKL coefficients
    (test/processnoise.jl:19-22).  The restated h-adaptive Genz-Malik driver (cubature.py, the host twin of
    b200e_koopman_solve) on the closed-form integrand u0 exp(1 + W8(1)) pdf(z): 401 nodes per region, and the
    expectation within the requested 1e-2 of prod_k M(a_k) (moment generating function of the +-4 sigma
    truncated normal)."""
    from scipy.special import ndtr
    a = np.array([math.sqrt(2.0) * math.sin((k + 0.5) * math.pi) / ((k + 0.5) * math.pi) for k in range(8)])
    Z = ndtr(4.0) - ndtr(-4.0)
    u0 = np.array([1.0, 2.0, 3.0, 4.0])

    def f(x):
        w = np.exp(-0.5 * np.sum(x * x, axis=1)) / ((2 * math.pi) ** 4 * Z ** 8)
        return (np.exp(1.0 + x @ a) * w)[:, None] * u0[None, :]

    rec = []
    r = sme.hcubature_v(f, [-4.0] * 8, [4.0] * 8, fdim=4, reltol=1e-2, abstol=1e-2, maxevals=12_000_000, record=rec)
    exact = u0 * math.e * np.prod(np.exp(a * a / 2) * (ndtr(4 - a) - ndtr(-4 - a)) / Z)
    assert len(rec[0]) == 401 and all(len(b) % 401 == 0 for b in rec)
    assert r.converged and np.max(np.abs(r.u - exact) / exact) < 1.1e-2
    assert np.all(r.resid >= np.abs(r.u - exact) * 0.5)      # the error estimate is not optimistic by more than 2x


def test_lamba_model_validation_and_compile(sme, tmp_path):
    spec = sme.models.builtin("gbm4_lamba", cache_dir=str(tmp_path))
    assert spec.solver == sme.models.SOLVER_LAMBA_EM
    nbytes, log = sme.compile_check(spec)
    assert nbytes > 10000 and "sm_100a" in log and "b200e_reduce_gen" in log
    for bad in (dict(n_kkl=0), dict(abstol=0.0, reltol=0.0), dict(abstol=-1.0), dict(save_times=(0.5,), nout=4)):
        with pytest.raises(sme.B200Error) as e:
            sme.compile_check(spec.replace(**bad))
        assert e.value.code == sme._capi.ERR_INVALID
    # the mirrored API lowers ProcessNoiseSystemMap(prob, n, LambaEM(); abstol, reltol) to that solver
    m = sme.models
    prob = sme.SDEProblem(m._GBM, "", [1.0, 2.0, 3.0, 4.0], (0.0, 1.0))
    sm = sme.ProcessNoiseSystemMap(prob, 8, sme.LambaEM(), abstol=1e-3, reltol=1e-3)
    ex = sme.ExpectationProblem(sm, m.observable_components([0, 1, 2, 3]), m.hmap_scatter(z_from_x=[(k, k) for k in range(8)]), nout=4)
    lowered = ex.model_spec()
    assert (lowered.solver, lowered.n_kkl, lowered.abstol, lowered.reltol) == (sme.models.SOLVER_LAMBA_EM, 8, 1e-3, 1e-3)
    with pytest.raises(NotImplementedError):
        sme.ExpectationProblem(sme.ProcessNoiseSystemMap(prob, 8), ex.g, ex.h, nout=4).model_spec()
    # the ensemble algorithm as the second forwarded solve argument, as in solve(ensprob, alg, ensemblealg; ...)
    sm2 = sme.ProcessNoiseSystemMap(prob, 8, sme.LambaEM(), sme.EnsembleB200(fp32=True, schedule=1), abstol=1e-3, reltol=1e-3)
    low2 = sme.ExpectationProblem(sm2, ex.g, ex.h, nout=4).model_spec()
    assert (low2.fp_mode, low2.schedule, low2.solver) == (1, 1, sme.models.SOLVER_LAMBA_EM)


# ---------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------
LAMBA = [("gbm4_lamba", {}), ("oscillator_noise", dict(solver=2, abstol=1e-4, reltol=1e-4, dt_fixed=0.0))]


def _spec_and_oracle(gpu, name, kw, **more):
    spec = gpu.models.builtin(name, **kw, **more)
    if name == "gbm4_lamba":
        orc = oracle.OracleProblem.builtin("gbm4", solver=oracle.SOLVER_LAMBA_EM, abstol=spec.abstol, reltol=spec.reltol)
    else:
        orc = oracle.OracleProblem.builtin(name, solver=oracle.SOLVER_LAMBA_EM, abstol=spec.abstol, reltol=spec.reltol)
    return spec, orc


@pytest.mark.gpu
@pytest.mark.parametrize("name,kw", LAMBA)
def test_lamba_nodes_match_oracle(gpu, name, kw):
    n = 3000
    x = sample_points("gbm4", n, seed=11)
    spec, orc = _spec_and_oracle(gpu, name, kw)
    ref, cref, sref = orc.eval_nodes(x, weight_by_pdf=True, want_steps=True)
    with gpu.Handle(spec) as h:
        dx, steps = h.eval_nodes(x, weight_by_pdf=True, want_steps=True)
        c = h.last_counters()
        st = h.last_stats()
    assert c == cref and np.array_equal(steps, sref)        # integer bookkeeping bit-exact
    scale = np.abs(ref).max(axis=0, keepdims=True)
    err = np.abs(dx - ref) / scale
    assert np.median(err) <= 1e-12 and err.max() <= 1e-8, (np.median(err), err.max())
    assert st["local_bytes_per_thread"] <= 64
    # the independent C model (oracle/models_builtin.c) and the twin compiled from the model's own source agree
    if name == "gbm4_lamba":
        dx2, c2 = oracle_for("gbm4_lamba").eval_nodes(x, weight_by_pdf=True)
        assert c2 == cref and np.allclose(dx2, ref, rtol=1e-12)


@pytest.mark.gpu
def test_lamba_golden_fixture(gpu):
    x = GOLD["gbm4_lamba_x"]
    with gpu.Handle(gpu.models.builtin("gbm4_lamba")) as h:
        dx, steps = h.eval_nodes(x, weight_by_pdf=True, want_steps=True)
        c = h.last_counters()
        s, s2, _ = h.reduce_samples(x)
    assert np.array_equal(steps, GOLD["gbm4_lamba_steps"])
    assert [c["nf"], c["naccept"], c["nreject"], c["nfail"]] == GOLD["gbm4_lamba_counters"].tolist()
    assert np.max(np.abs(dx - GOLD["gbm4_lamba_dx"]) / np.abs(GOLD["gbm4_lamba_dx"]).max(axis=0)) <= 1e-8
    assert np.allclose(s, GOLD["gbm4_lamba_sum"], rtol=1e-8) and np.allclose(s2, GOLD["gbm4_lamba_sumsq"], rtol=1e-8)


@pytest.mark.gpu
@pytest.mark.parametrize("name,kw", LAMBA)
@pytest.mark.parametrize("schedule", [0, 1, 2])
def test_lamba_reduce_matches_oracle_under_every_schedule(gpu, name, kw, schedule):
    n = 20011
    x = sample_points("gbm4", n, seed=12)
    spec, orc = _spec_and_oracle(gpu, name, kw, schedule=schedule)
    so, s2o, co = orc.reduce_samples(x)
    with gpu.Handle(spec) as h:
        s, s2, c = h.reduce_samples(x)
        sb, s2b, cb = h.reduce_samples(x)
    assert c == co == cb
    assert np.max(np.abs(s - so) / np.abs(so)) < 1e-8 and np.max(np.abs(s2 - s2o) / np.abs(s2o)) < 1e-8
    if schedule:  # static ownership / super-chunk rows: bit-reproducible sums
        assert np.array_equal(s, sb) and np.array_equal(s2, s2b)


@pytest.mark.gpu
def test_lamba_sampler_in_the_kernel_and_closed_form(gpu):
    """MonteCarlo with rand(d) drawn in the kernel: same sums as the oracle on the host twin's samples, and
    the expectation against E[u0 exp(1 + W8(1))] = u0 e prod_k M(a_k), M the moment generating function of
    the +-4 sigma truncated normal -- within the scheme's own bias at tol 1e-3 (test/processnoise.jl:25 asks
    1e-2 between two runs of the same scheme)."""
    from scipy.special import ndtr
    n = 200000
    spec, orc = _spec_and_oracle(gpu, "gbm4_lamba", {})
    x = gpu.sample_host(spec, seed=21, first_index=0, n=n)
    so, s2o, co = orc.reduce_samples(x)
    with gpu.Handle(spec) as h:
        s, s2, c = h.reduce_generated(seed=21, n=n)
    assert c == co and np.max(np.abs(s - so) / np.abs(so)) < 1e-8 and np.max(np.abs(s2 - s2o) / np.abs(s2o)) < 1e-8
    a = np.array([math.sqrt(2.0) * math.sin((k + 0.5) * math.pi) / ((k + 0.5) * math.pi) for k in range(8)])
    mgf = np.exp(a * a / 2) * (ndtr(4 - a) - ndtr(-4 - a)) / (ndtr(4.0) - ndtr(-4.0))
    exact = np.array([1.0, 2.0, 3.0, 4.0]) * math.e * np.prod(mgf)
    assert np.allclose(s / n, exact, rtol=0.05), (s / n, exact)
    # tighter tolerances move the expectation toward the closed form
    with gpu.Handle(spec.replace(abstol=1e-5, reltol=1e-5)) as h:
        st, _, ct = h.reduce_generated(seed=21, n=n)
    assert ct["nfail"] == 0 and np.max(np.abs(st / n - exact) / exact) < 0.4 * np.max(np.abs(s / n - exact) / exact) + 0.01


@pytest.mark.gpu
def test_lamba_through_the_mirrored_api_and_fp32(gpu):
    """test/processnoise.jl:9-22 through the mirrored API; FP32 state arithmetic stays within its own rounding."""
    m = gpu.models
    prob = gpu.SDEProblem(m._GBM, "", [1.0, 2.0, 3.0, 4.0], (0.0, 1.0))
    sm = gpu.ProcessNoiseSystemMap(prob, 8, gpu.LambaEM(), abstol=1e-3, reltol=1e-3)
    ex = gpu.ExpectationProblem(sm, m.observable_components([0, 1, 2, 3]), m.hmap_scatter(z_from_x=[(k, k) for k in range(8)]), nout=4)
    sol = gpu.solve(ex, gpu.MonteCarlo(100000, seed=5))
    spec, orc = _spec_and_oracle(gpu, "gbm4_lamba", {})
    x = gpu.sample_host(spec, seed=5, first_index=0, n=100000)
    so, _, _ = orc.reduce_samples(x)
    assert np.allclose(sol.u, so / 100000, rtol=1e-8)
    sm32 = gpu.ProcessNoiseSystemMap(prob, 8, gpu.LambaEM(), abstol=1e-3, reltol=1e-3, ensemblealg=gpu.EnsembleB200(fp32=True))
    ex32 = gpu.ExpectationProblem(sm32, ex.g, ex.h, nout=4)
    sol32 = gpu.solve(ex32, gpu.MonteCarlo(100000, seed=5))
    assert np.allclose(sol32.u, sol.u, rtol=2e-3)
    ex.close(); ex32.close()


@pytest.mark.gpu
def test_reference_processnoise_testset_koopman_vs_montecarlo(gpu):
    """reference test/processnoise.jl:9-26 end to end: Koopman (ireltol = iabstol = 1e-3) against
    MonteCarlo(1_000_000), both with LambaEM(abstol = reltol = 1e-3), must agree at rtol = 1e-2.  The reference
    integrates with CubaDivonne; here the 8-dimensional integral goes through the h-adaptive Genz-Malik rule on
    the device (401 nodes per region), which needs more than the reference's default 1e6 evaluations."""
    m = gpu.models
    prob = gpu.SDEProblem(m._GBM, "", [1.0, 2.0, 3.0, 4.0], (0.0, 1.0))
    sm = gpu.ProcessNoiseSystemMap(prob, 8, gpu.LambaEM(), abstol=1e-3, reltol=1e-3)
    ex = gpu.ExpectationProblem(sm, m.observable_components([0, 1, 2, 3]), m.hmap_scatter(z_from_x=[(k, k) for k in range(8)]), nout=4)
    sol1 = gpu.solve(ex, gpu.Koopman(), ireltol=1e-3, iabstol=1e-3, batch=64, quadalg=gpu.B200Cubature(), maxiters=40_000_000)
    sol2 = gpu.solve(ex, gpu.MonteCarlo(1_000_000))
    assert np.allclose(sol1.u, sol2.u, rtol=1e-2), (sol1.u, sol2.u)     # the reference's own assertion
    st = sol1.original
    assert st["counters"]["nfail"] == 0 and st["nevals"] >= 1_000_000
    ex.close()


@pytest.mark.gpu
def test_lamba_dtmax_and_maxiters_are_honoured(gpu):
    """solve kwargs forwarded through ProcessNoiseSystemMap (src/system_utils.jl:98-100): a dtmax below the span
    clips every proposed step, maxiters aborts are counted per trajectory (g still applied to the last state)."""
    n = 2000
    x = sample_points("gbm4", n, seed=13)
    for kw in (dict(dtmax=0.004), dict(maxiters=60)):
        spec = gpu.models.builtin("gbm4_lamba", **kw)
        orc = oracle.OracleProblem.builtin("gbm4", solver=oracle.SOLVER_LAMBA_EM, abstol=1e-3, reltol=1e-3, **kw)
        ref, cref, sref = orc.eval_nodes(x, weight_by_pdf=False, want_steps=True)
        with gpu.Handle(spec) as h:
            dx, steps = h.eval_nodes(x, weight_by_pdf=False, want_steps=True)
            c = h.last_counters()
        assert c == cref and np.array_equal(steps, sref), kw
        assert np.max(np.abs(dx - ref) / np.abs(ref).max(axis=0)) <= 1e-8
        if "maxiters" in kw:
            assert c["nfail"] == n and int(steps.max()) == 60
        else:
            assert c["nfail"] == 0 and int(steps.min()) >= 250      # at least (t1 - t0) / dtmax steps
