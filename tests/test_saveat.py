"""Time-sampled observables: ``g(sol,p) = sol(t_k)`` / ``sol[i, :]`` with ``saveat``
(reference docs/src/tutorials/introduction.md:53, :192, :207-212).  CPU part: the oracle's restatement of
Tsit5's free interpolant against closed forms; the GPU part compares the kernel with the oracle."""
import numpy as np
import pytest
from scipy.linalg import expm

from tests.sampling_util import oracle_for, sample_points

A = np.array([[0.0, 1.0], [-1.0, -2.0]])
TS6 = (0.0, 0.5, 1.0, 1.7, 2.5, 3.0)


def test_tsit5_interpolant_order_conditions():
    """b_i(theta) of the oracle / kernel satisfy the quadrature and tree conditions through order 4 for
    every theta, and reduce to the step's own weights at theta = 1."""
    c = np.array([0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0, 1.0])
    a = np.zeros((7, 7))
    a[1, 0] = 0.161
    a[2, :2] = [-0.008480655492356989, 0.335480655492357]
    a[3, :3] = [2.8971530571054935, -6.359448489975075, 4.3622954328695815]
    a[4, :4] = [5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525]
    a[5, :5] = [5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383]
    a[6, :6] = [0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774]
    r = np.array([[1.0, -2.763706197274826, 2.9132554618219126, -1.0530884977290216],
                  [0, 0.13169999999999998, -0.2234, 0.1017],
                  [0, 3.9302962368947516, -5.941033872131505, 2.490627285651253],
                  [0, -12.411077166933676, 30.33818863028232, -16.548102889244902],
                  [0, 37.50931341651104, -88.1789048947664, 47.37952196281928],
                  [0, -27.896526289197286, 65.09189467479366, -34.87065786149661],
                  [0, 1.5, -4.0, 2.5]])
    b = lambda th: np.array([th * (ri[0] + th * (ri[1] + th * (ri[2] + th * ri[3]))) for ri in r])
    assert np.max(np.abs(b(1.0) - np.append(a[6, :6], 0.0))) < 1e-14
    for th in (0.05, 0.3, 0.5, 0.77, 0.99):
        bt = b(th)
        for lhs, rhs in ((bt.sum(), th), (bt @ c, th ** 2 / 2), (bt @ c ** 2, th ** 3 / 3), (bt @ c ** 3, th ** 4 / 4),
                         (bt @ (a @ c), th ** 3 / 6), (bt @ (a @ c ** 2), th ** 4 / 12), (bt @ (a @ (a @ c)), th ** 4 / 24),
                         (bt @ (c * (a @ c)), th ** 4 / 8)):
            assert abs(lhs - rhs) < 1e-13


def test_oracle_saveat_matches_closed_form():
    x = sample_points("linear2_saveat", 500, seed=21)
    ref = np.stack([x @ expm(A * t).T for t in TS6], axis=1).reshape(len(x), -1)
    # tight tolerances: the 4th-order interpolant converges to the exact solution
    dx, c = oracle_for("linear2_saveat", abstol=1e-11, reltol=1e-11).eval_nodes(x, weight_by_pdf=False)
    assert np.max(np.abs(dx - ref)) < 1e-9 and c["nfail"] == 0
    # default tolerances: error at the level of the solver tolerance, and the slots at t0 / t1 are the
    # initial state / the end state of the plain end-state problem bit for bit
    dx, c = oracle_for("linear2_saveat").eval_nodes(x, weight_by_pdf=False)
    assert np.max(np.abs(dx - ref)) < 2e-4
    assert np.array_equal(dx[:, :2], x)
    end, cend = oracle_for("linear2").eval_nodes(x, weight_by_pdf=False)
    assert np.array_equal(dx[:, -2:], end) and c == cend      # saving does not change the stepping


def test_oracle_reproduces_the_tutorial_expectation():
    """introduction.md:207-212: E[u(t)] = exp(p t) mean(u0) for u' = p u, u0 ~ Uniform(0, 10), p = -0.3."""
    x = sample_points("growth_saveat", 200000, seed=22)
    s, s2, c = oracle_for("growth_saveat").reduce_samples(x)
    ts = 0.5 * np.arange(21)
    assert np.allclose(s / len(x), np.exp(-0.3 * ts) * x.mean(), rtol=2e-4)
    assert np.allclose(s / len(x), np.exp(-0.3 * ts) * 5.0, rtol=1e-2)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["linear2_saveat", "growth_saveat"])
def test_gpu_saveat_matches_oracle(gpu, name):
    n = 50000
    x = sample_points(name, n, seed=23)
    orc = oracle_for(name)
    ref, cref, sref = orc.eval_nodes(x, weight_by_pdf=True, want_steps=True)
    so, s2o, co = orc.reduce_samples(x)
    with gpu.Handle(gpu.models.builtin(name)) as h:
        dx, steps = h.eval_nodes(x, weight_by_pdf=True, want_steps=True)
        c = h.last_counters()
        s, s2, cr = h.reduce_samples(x)
        sg, s2g, cg = h.reduce_generated(9, n)
        xg = h.sample_host(9, n)
    assert c == cref == cr and np.array_equal(steps, sref)
    scale = np.abs(ref).max(axis=0, keepdims=True) + 1e-300
    err = np.abs(dx - ref) / scale
    assert np.median(err) <= 1e-10 and err.max() <= 2e-6
    assert np.max(np.abs(s - so) / np.abs(so)) < 1e-8 and np.max(np.abs(s2 - s2o) / np.abs(s2o)) < 1e-8
    sgo, _, cgo = orc.reduce_samples(xg)
    assert cg == cgo and np.max(np.abs(sg - sgo) / np.abs(sgo)) < 1e-8


@pytest.mark.gpu
def test_gpu_tutorial_expectation_through_the_api(gpu):
    """introduction.md:207-212 through the mirrored public API: Koopman and MonteCarlo of sol[1, :]."""
    m = gpu.models
    ts = [0.5 * k for k in range(21)]
    prob = gpu.ODEProblem(m._SCALAR_GROWTH_RHS, [10.0], (0.0, 10.0), [-0.3], saveat=ts)
    smap = gpu.SystemMap(prob, gpu.Tsit5(), gpu.EnsembleB200())
    gd = gpu.GenericDistribution(gpu.Uniform(0.0, 10.0))
    ex = gpu.ExpectationProblem(smap, m.observable_at_components([0]), m.hmap_scatter(u0_from_x=[(0, 0)]), gd, nout=21)
    exact = 5.0 * np.exp(-0.3 * np.array(ts))
    sol = gpu.solve(ex, gpu.Koopman(), batch=1000, quadalg=gpu.B200Cubature(), ireltol=1e-6, iabstol=1e-6)
    assert np.allclose(sol.u, exact, rtol=2e-4)
    mc = gpu.solve(ex, gpu.MonteCarlo(400000, seed=4))
    assert np.allclose(mc.u, exact, rtol=1e-2)
    ex.close()
