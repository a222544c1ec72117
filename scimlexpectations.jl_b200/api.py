"""Host-side mirror of the reference's public interface for the batched-integrand path.

Same names, argument order and semantics as SciMLExpectations.jl (the Julia toolchain is absent from
this image, so the host side above the C ABI is Python here and a ``ccall`` shim in
``julia/B200Expectations.jl``):

    GenericDistribution(dists...)                      src/distribution_utils.jl:33-54
    SystemMap(prob, alg, ensemblealg; kwargs...)       src/system_utils.jl:32-62
    ProcessNoiseSystemMap(prob, n, args...; kwargs...) src/system_utils.jl:92-117
    ExpectationProblem(sm, g, h, d)                    src/problem_types.jl:41-85
    solve(exprob, Koopman(); batch, quadalg, ireltol, iabstol, maxiters)   src/expectation.jl:336-354
    solve(exprob, MonteCarlo(n))                       src/expectation.jl:268-325
    build_integrand(exprob, Koopman(), mid, p, batch)  src/expectation.jl:169-200, :209-249
    ExpectationSolution(u, resid, original)            src/solution_types.jl:14-18

What changes relative to the reference is only what sits in the ``ensemblealg`` slot:
``EnsembleB200()``.  Arbitrary host closures cannot run on the device, so ``f`` (the ODE right-hand
side), ``g`` (observable) and ``h`` (covariate map) are given as CUDA device-function source in the
dialect of ``csrc/kernels/expect_prelude.cuh`` -- or built from the structured helpers
``models.hmap_scatter`` / ``models.observable_components``.  There is no CPU fallback: without the
CUDA library and a B200 ``solve`` raises.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Any, Optional, Sequence

import numpy as np
from scipy.special import ndtr, ndtri

from . import _capi
from ._capi import (FP32, FP64, LAYOUT_POINT_MAJOR, ModelSpec, PDF_LOGPOLY, PDF_NORMAL, PDF_TABLE, PDF_TRUNCNORMAL,
                    PDF_UNIFORM)
from .cubature import hcubature_v

# ---------------------------------------------------------------------------------------------
# distributions (the subset of Distributions.jl the product density table covers)
# ---------------------------------------------------------------------------------------------


@dataclass(frozen=True)
class Uniform:
    a: float = 0.0
    b: float = 1.0

    def minimum(self): return float(self.a)
    def maximum(self): return float(self.b)
    def logpdf(self, x): return -math.log(self.b - self.a) if self.a <= x <= self.b else -math.inf
    def icdf(self, u): return self.a + (self.b - self.a) * u
    def table(self): return (PDF_UNIFORM, self.a, self.b, 0.0, 1.0)


@dataclass(frozen=True)
class Normal:
    mu: float = 0.0
    sigma: float = 1.0

    def minimum(self): return -math.inf
    def maximum(self): return math.inf

    def logpdf(self, x):
        z = (x - self.mu) / self.sigma
        return -0.5 * z * z - math.log(self.sigma) - 0.5 * math.log(2.0 * math.pi)

    def icdf(self, u): return self.mu + self.sigma * ndtri(u)
    def table(self): return (PDF_NORMAL, -math.inf, math.inf, self.mu, self.sigma)


@dataclass(frozen=True)
class Truncated:
    """``truncated(Normal(mu, sigma), lower, upper)``"""
    untruncated: Normal
    lower: float
    upper: float

    def minimum(self): return float(self.lower)
    def maximum(self): return float(self.upper)

    def _tp(self):
        d = self.untruncated
        return float(ndtr((self.upper - d.mu) / d.sigma) - ndtr((self.lower - d.mu) / d.sigma))

    def logpdf(self, x):
        if not (self.lower <= x <= self.upper):
            return -math.inf
        return self.untruncated.logpdf(x) - math.log(self._tp())

    def icdf(self, u):
        d = self.untruncated
        a, b = ndtr((self.lower - d.mu) / d.sigma), ndtr((self.upper - d.mu) / d.sigma)
        return np.clip(d.mu + d.sigma * ndtri(a + (b - a) * u), self.lower, self.upper)

    def table(self):
        return (PDF_TRUNCNORMAL, self.lower, self.upper, self.untruncated.mu, self.untruncated.sigma)


class _LogPoly:
    """A factor whose log-density is ``c0 + c1 x + c2 x^2 + c3 log x + c4 log(1 - x) + c5 log(x)^2`` on its support
    (``B200E_PDF_LOGPOLY``): the exponential-family members of Distributions.jl a prior is usually built from.  The
    device has no inverse CDF for them: ``MonteCarlo`` ships host samples (``scipy.stats`` quantiles of the block-keyed
    uniform stream) for a density that holds one."""
    device_sampler = False
    _frozen = None   # scipy.stats frozen distribution

    def coefficients(self):   # (c0, c1, c2, c3, c4, c5) on the untruncated support
        raise NotImplementedError

    def minimum(self): return float(self._frozen.support()[0])
    def maximum(self): return float(self._frozen.support()[1])

    def logpdf(self, x):
        if not (self.minimum() <= x <= self.maximum()):
            return -math.inf
        c = self.coefficients()
        lp = c[0] + c[1] * x + c[2] * x * x
        if c[3] or c[5]:
            lx = math.log(x) if x > 0 else -math.inf
            lp += (c[5] * lx + c[3]) * lx if c[5] else c[3] * lx
        if c[4]:
            lp += c[4] * math.log1p(-x) if x < 1 else -math.inf
        return lp

    def cdf(self, x): return float(self._frozen.cdf(x))
    def icdf(self, u): return self._frozen.ppf(u)
    def table(self): return (PDF_LOGPOLY, self.minimum(), self.maximum(), 0.0, 1.0, np.asarray(self.coefficients(), dtype=np.float64))


class Exponential(_LogPoly):
    """``Exponential(theta)`` (scale): logpdf = -log(theta) - x / theta on [0, inf)."""
    def __init__(self, theta: float = 1.0):
        from scipy import stats
        self.theta = float(theta)
        self._frozen = stats.expon(scale=self.theta)

    def coefficients(self): return (-math.log(self.theta), -1.0 / self.theta, 0.0, 0.0, 0.0, 0.0)


class Gamma(_LogPoly):
    """``Gamma(alpha, theta)`` (shape, scale): logpdf = (alpha - 1) log x - x / theta - lgamma(alpha) - alpha log(theta)."""
    def __init__(self, alpha: float = 1.0, theta: float = 1.0):
        from scipy import stats
        self.alpha, self.theta = float(alpha), float(theta)
        self._frozen = stats.gamma(self.alpha, scale=self.theta)

    def coefficients(self):
        return (-math.lgamma(self.alpha) - self.alpha * math.log(self.theta), -1.0 / self.theta, 0.0, self.alpha - 1.0, 0.0, 0.0)


class Beta(_LogPoly):
    """``Beta(alpha, beta)``: logpdf = (alpha - 1) log x + (beta - 1) log(1 - x) - lbeta(alpha, beta) on [0, 1]."""
    def __init__(self, alpha: float = 1.0, beta: float = 1.0):
        from scipy import stats
        self.alpha, self.beta = float(alpha), float(beta)
        self._frozen = stats.beta(self.alpha, self.beta)

    def coefficients(self):
        lbeta = math.lgamma(self.alpha) + math.lgamma(self.beta) - math.lgamma(self.alpha + self.beta)
        return (-lbeta, 0.0, 0.0, self.alpha - 1.0, self.beta - 1.0, 0.0)


class LogNormal(_LogPoly):
    """``LogNormal(mu, sigma)``: logpdf = -log x - log(sigma) - log(2 pi)/2 - (log x - mu)^2 / (2 sigma^2) on (0, inf)."""
    def __init__(self, mu: float = 0.0, sigma: float = 1.0):
        from scipy import stats
        self.mu, self.sigma = float(mu), float(sigma)
        self._frozen = stats.lognorm(self.sigma, scale=math.exp(self.mu))

    def coefficients(self):
        s2 = self.sigma * self.sigma
        return (-math.log(self.sigma) - 0.5 * math.log(2.0 * math.pi) - self.mu * self.mu / (2.0 * s2), 0.0, 0.0,
                self.mu / s2 - 1.0, 0.0, -1.0 / (2.0 * s2))


class Rayleigh(_LogPoly):
    """``Rayleigh(sigma)``: logpdf = log x - x^2 / (2 sigma^2) - 2 log(sigma) on [0, inf)."""
    def __init__(self, sigma: float = 1.0):
        from scipy import stats
        self.sigma = float(sigma)
        self._frozen = stats.rayleigh(scale=self.sigma)

    def coefficients(self): return (-2.0 * math.log(self.sigma), 0.0, -0.5 / (self.sigma * self.sigma), 1.0, 0.0, 0.0)


class Chisq(_LogPoly):
    """``Chisq(nu)`` = Gamma(nu / 2, 2)."""
    def __init__(self, nu: float):
        from scipy import stats
        self.nu = float(nu)
        self._frozen = stats.chi2(self.nu)

    def coefficients(self):
        h = 0.5 * self.nu
        return (-math.lgamma(h) - h * math.log(2.0), -0.5, 0.0, h - 1.0, 0.0, 0.0)


class Chi(_LogPoly):
    """``Chi(nu)``: logpdf = (nu - 1) log x - x^2 / 2 - (nu / 2 - 1) log 2 - lgamma(nu / 2) on [0, inf)."""
    def __init__(self, nu: float):
        from scipy import stats
        self.nu = float(nu)
        self._frozen = stats.chi(self.nu)

    def coefficients(self):
        h = 0.5 * self.nu
        return (-(h - 1.0) * math.log(2.0) - math.lgamma(h), 0.0, -0.5, self.nu - 1.0, 0.0, 0.0)


class Erlang(Gamma):
    """``Erlang(k, theta)`` = Gamma with an integer shape."""
    def __init__(self, k: int = 1, theta: float = 1.0):
        if int(k) != k or k < 1:
            raise ValueError("Erlang takes a positive integer shape")
        super().__init__(float(k), theta)


class Pareto(_LogPoly):
    """``Pareto(alpha, theta)`` (shape, scale): logpdf = log(alpha) + alpha log(theta) - (alpha + 1) log x on [theta, inf)."""
    def __init__(self, alpha: float = 1.0, theta: float = 1.0):
        from scipy import stats
        self.alpha, self.theta = float(alpha), float(theta)
        self._frozen = stats.pareto(self.alpha, scale=self.theta)

    def coefficients(self):
        return (math.log(self.alpha) + self.alpha * math.log(self.theta), 0.0, 0.0, -(self.alpha + 1.0), 0.0, 0.0)


class TruncatedLogPoly(_LogPoly):
    """``truncated(d, lower, upper)`` of one of the above: the same coefficients with ``-log(cdf(upper) - cdf(lower))``
    folded into ``c0`` (Distributions.logpdf of a Truncated), support ``[lower, upper]``."""
    def __init__(self, d: _LogPoly, lower: float, upper: float):
        self.untruncated, self.lower, self.upper = d, max(float(lower), d.minimum()), min(float(upper), d.maximum())
        if not self.upper > self.lower:
            raise ValueError("empty truncation interval")
        self._a, self._b = d.cdf(self.lower), d.cdf(self.upper)

    def minimum(self): return self.lower
    def maximum(self): return self.upper

    def coefficients(self):
        c = list(self.untruncated.coefficients())
        c[0] -= math.log(self._b - self._a)
        return tuple(c)

    def cdf(self, x): return (self.untruncated.cdf(min(max(x, self.lower), self.upper)) - self._a) / (self._b - self._a)
    def icdf(self, u): return np.clip(self.untruncated.icdf(self._a + (self._b - self._a) * np.asarray(u)), self.lower, self.upper)


class UnivariateKDE:
    """A density tabulated on a uniform grid -- what ``kde(samples)`` of KernelDensity.jl returns (fields ``x``,
    ``density``) and what ``pdf(kde, x)`` interpolates (docs/src/tutorials/gpu_bayesian.md:153-166:
    ``pdf_func(x) = prod(pdf.(p_kde, x))``, ``lb / ub`` = the grid ends).  Quadratic B-spline through the grid
    values (KernelDensity's ``InterpKDE`` default), 0 outside.  Koopman only (``rand_func() = nothing`` in the tutorial): it cannot be sampled."""

    def __init__(self, x, density):
        self.x = np.ascontiguousarray(x, dtype=np.float64)
        self.density = np.ascontiguousarray(density, dtype=np.float64)
        if self.x.ndim != 1 or self.x.shape != self.density.shape or len(self.x) < 2:
            raise ValueError("x and density must be 1-D arrays of equal length >= 2")
        if not np.allclose(np.diff(self.x), (self.x[-1] - self.x[0]) / (len(self.x) - 1), rtol=1e-9, atol=0):
            raise ValueError("the grid must be uniform")

    @classmethod
    def from_samples(cls, samples, npoints: int = 2048, bandwidth: Optional[float] = None):
        """Gaussian KDE on ``npoints`` grid points with Silverman's rule and a 4-bandwidth margin (the defaults
        of KernelDensity.jl's ``kde``)."""
        v = np.asarray(samples, dtype=np.float64).ravel()
        if bandwidth is None:
            q75, q25 = np.percentile(v, [75, 25])
            width = min(v.std(ddof=1), (q75 - q25) / 1.34)
            bandwidth = 0.9 * width * len(v) ** (-0.2)
        lo, hi = v.min() - 4.0 * bandwidth, v.max() + 4.0 * bandwidth
        x = np.linspace(lo, hi, npoints)
        z = (x[:, None] - v[None, :]) / bandwidth
        dens = np.exp(-0.5 * z * z).sum(axis=1) / (len(v) * bandwidth * math.sqrt(2.0 * math.pi))
        return cls(x, dens)

    def minimum(self): return float(self.x[0])
    def maximum(self): return float(self.x[-1])

    def spline_coefficients(self) -> np.ndarray:
        """The ``n + 2`` coefficients of the quadratic B-spline through the grid values with zero second derivative
        in the end cells -- Interpolations' ``BSpline(Quadratic(Line(OnGrid())))``, what KernelDensity's
        ``pdf(kde, x)`` evaluates (the library does the same prefilter in ``b200e_create``)."""
        if getattr(self, "_coef", None) is None:
            from scipy.linalg import solve_banded
            f, n = self.density, len(self.density)
            c = np.empty(n + 2)
            c[1], c[n] = f[0], f[-1]
            if n > 2:
                ab = np.zeros((3, n - 2))
                ab[0, 1:], ab[1, :], ab[2, :-1] = 0.125, 0.75, 0.125
                rhs = f[1:-1].copy()
                rhs[0] -= 0.125 * c[1]
                rhs[-1] -= 0.125 * c[n]
                c[2:n] = solve_banded((1, 1), ab, rhs)
            c[0], c[n + 1] = 2 * c[1] - c[2], 2 * c[n] - c[n - 1]
            self._coef = c
        return self._coef

    def pdf(self, x):
        """``pdf(kde, x)``: quadratic B-spline interpolation of the grid values, 0 outside the grid."""
        x = float(x)
        if not (self.x[0] <= x <= self.x[-1]):
            return 0.0
        n, c = len(self.x), self.spline_coefficients()
        g = (x - self.x[0]) * (n - 1) / (self.x[-1] - self.x[0])
        i = min(int(g + 0.5), n - 1)
        d = g - i
        return float(0.5 * (d - 0.5) ** 2 * c[i] + (0.75 - d * d) * c[i + 1] + 0.5 * (d + 0.5) ** 2 * c[i + 2])

    def logpdf(self, x):
        v = self.pdf(x)
        return math.log(v) if v > 0 else -math.inf

    def icdf(self, u):
        raise NotImplementedError("a tabulated density cannot be sampled (Koopman only)")

    def table(self):
        return (PDF_TABLE, float(self.x[0]), float(self.x[-1]), 0.0, 1.0, self.density)


def truncated(d, lower: float, upper: float):
    """``truncated(d, lower, upper)``: Normal (its own device kind, which the in-kernel sampler can draw) and the
    log-polynomial factors (Exponential, Gamma, Beta, LogNormal, Rayleigh, Chisq, Chi, Erlang, Pareto)."""
    if isinstance(d, Normal):
        return Truncated(d, float(lower), float(upper))
    if isinstance(d, _LogPoly):
        return TruncatedLogPoly(d, lower, upper)
    raise TypeError("truncated() takes a Normal or a log-polynomial factor (Exponential, Gamma, Beta, LogNormal, Rayleigh, Chisq, "
                    "Chi, Erlang, Pareto)")


class GenericDistribution:
    """Product of independent univariate distributions (reference distribution_utils.jl:40-48):
    ``pdf_func(x) = exp(sum(logpdf(f, y)))``, ``rand_func() = [rand(d) for d in dists]``,
    ``lb/ub = minimum/maximum`` of the factors."""

    def __init__(self, d, *ds):
        self.dists = (d,) + tuple(ds)
        self.lb = np.array([x.minimum() for x in self.dists], dtype=np.float64)
        self.ub = np.array([x.maximum() for x in self.dists], dtype=np.float64)
        self._rng = np.random.default_rng()

    def pdf_func(self, x):
        return math.exp(sum(f.logpdf(float(y)) for f, y in zip(self.dists, x)))

    def rand_func(self):
        return np.array([d.icdf(self._rng.random()) for d in self.dists])

    def table(self):
        return [d.table() for d in self.dists]

    def __len__(self):
        return len(self.dists)

    # host-generated sample sets: block b of the stream is a pure function of (seed, b)
    def rand_block(self, seed: int, block: int, n: int) -> np.ndarray:
        rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence([int(seed), int(block)])))
        u = rng.random((n, len(self.dists)))
        out = np.empty_like(u)
        for j, d in enumerate(self.dists):
            out[:, j] = d.icdf(u[:, j])
        return out


def pdf(d: GenericDistribution, x): return d.pdf_func(x)
def minimum(d: GenericDistribution): return d.lb
def maximum(d: GenericDistribution): return d.ub
def extrema(d: GenericDistribution): return d.lb, d.ub
def rand(d: GenericDistribution): return d.rand_func()


# ---------------------------------------------------------------------------------------------
# problems, solvers, ensemble algorithm
# ---------------------------------------------------------------------------------------------
@dataclass
class ODEProblem:
    """``ODEProblem(f, u0, tspan, p; saveat)``; ``f`` is CUDA source defining ``rhs(du,u,p,t)``.
    ``saveat`` (docs/src/tutorials/introduction.md:207-212) lists the times at which the observable sees
    the solution: ``g`` is then an ``observable_at`` source and ``nout = len(saveat) * outputs per time``."""
    f: str
    u0: Sequence[float]
    tspan: Sequence[float]
    p: Sequence[float] = ()
    saveat: Sequence[float] = ()


@dataclass
class SDEProblem:
    """``SDEProblem(f, g, u0, tspan, p)`` with scalar noise; ``f``/``g`` are CUDA sources defining
    ``rhs`` and ``diffusion`` (the scalar dW multiplies every component, test/processnoise.jl:9-14)."""
    f: str
    g: str
    u0: Sequence[float]
    tspan: Sequence[float]
    p: Sequence[float] = ()


@dataclass(frozen=True)
class Tsit5:
    pass


@dataclass(frozen=True)
class EM:
    """Fixed-step Euler-Maruyama; ``dt`` may also be passed as the solve kwarg ``dt=``."""
    dt: float = 0.0


@dataclass(frozen=True)
class LambaEM:
    """Adaptive Euler-Maruyama with StochasticDiffEq's defaults (test/processnoise.jl:15); tolerances come
    from the ``abstol`` / ``reltol`` kwargs of ``ProcessNoiseSystemMap`` (SciML defaults 1e-2 / 1e-2 for SDEs)."""
    pass


@dataclass
class EnsembleB200:
    """The ``ensemblealg`` that routes the ensemble solve to libb200expect (``<: EnsembleAlgorithm``
    in the Julia shim).  Launch-shape fields are 0 for the library defaults."""
    devices: Optional[Sequence[int]] = None
    fp32: bool = False
    block_size: int = 0
    blocks_per_sm: int = 0
    refill_threshold: int = 0
    chunk_points: int = 0
    schedule: int = 0   # 0 dynamic work distribution (default), 1 static, 2 dynamic with bit-reproducible sums


class SystemMap:
    def __init__(self, prob: ODEProblem, alg: Optional[Tsit5] = None, ensemblealg: Optional[EnsembleB200] = None,
                 **kwargs):
        self.prob = prob
        self.alg = alg
        self.ensemblealg = ensemblealg if ensemblealg is not None else EnsembleB200()
        self.kwargs = kwargs  # abstol, reltol, dtmax, maxiters, save_everystep (ignored: end-state only)


class ProcessNoiseSystemMap:
    def __init__(self, prob: SDEProblem, n: int, *args, **kwargs):
        self.prob = prob
        self.n = int(n)
        self.args = args      # (EM(dt),) or (LambaEM(),), optionally followed by EnsembleB200(...) -- forwarded solve arguments
        self.kwargs = kwargs  # dt=... / abstol=..., reltol=..., ensemblealg=EnsembleB200(...)


class ExpectationProblem:
    """``ExpectationProblem(sm::SystemMap, g, h, d)`` / ``ExpectationProblem(sm::ProcessNoiseSystemMap, g, h)``.

    ``g``: CUDA source defining ``observable(u_end, p, t_end, out)`` (``g(sol,p)`` restricted to
    end-state observables); ``h``: CUDA source defining ``hmap``; ``nout`` is the observable's length
    (the reference learns it from a prototype solve, here the kernel needs it at compile time)."""

    def __init__(self, S, g: str, h: str, d: Optional[GenericDistribution] = None, *, nout: int = 1):
        self.S, self.g, self.h = S, g, h
        if isinstance(S, ProcessNoiseSystemMap):
            if d is not None:
                raise TypeError("ExpectationProblem(sm::ProcessNoiseSystemMap, g, h) takes no distribution")
            # problem_types.jl:83: n x Truncated(Normal(), -4, 4)
            d = GenericDistribution(*[truncated(Normal(), -4.0, 4.0) for _ in range(S.n)])
            self.params = np.array(S.prob.p, dtype=np.float64).copy()
        else:
            if d is None:
                raise TypeError("ExpectationProblem(sm::SystemMap, g, h, d) needs a distribution")
            self.params = (np.array(S.prob.u0, dtype=np.float64).copy(), np.array(S.prob.p, dtype=np.float64).copy())
        self.d = d
        self.nout = int(nout)
        self._handle = None
        self._handle_key = None

    # getters of the reference (problem_types.jl:68-72)
    def distribution(self): return self.d
    def mapping(self): return self.S
    def observable(self): return self.g
    def input_cov(self): return self.h
    def parameters(self): return self.params

    # ---- lowering to the C ABI ----
    def model_spec(self) -> ModelSpec:
        S = self.S
        prob = S.prob
        if isinstance(S, ProcessNoiseSystemMap):
            # the reference forwards args... to solve(ensprob, alg, ensemblealg; ...): the ensemble algorithm is positional there
            ens = next((a for a in S.args if isinstance(a, EnsembleB200)), S.kwargs.get("ensemblealg", EnsembleB200()))
            alg = next((a for a in S.args if isinstance(a, (EM, LambaEM))), None)
            if alg is None:
                raise NotImplementedError("the B200 process-noise path implements EM(dt) and LambaEM(); pass one in args")
            adaptive = isinstance(alg, LambaEM)
            dt = 0.0 if adaptive else float(S.kwargs.get("dt", alg.dt))
            if not adaptive and not dt > 0:
                raise ValueError("EM needs dt > 0")
            src = prob.f + "\n" + prob.g + "\n" + self.h + "\n" + self.g
            return ModelSpec(source=src, nstate=len(prob.u0), nparam=len(prob.p), nx=len(self.d), nout=self.nout,
                             u0=list(prob.u0), p=list(prob.p), tspan=tuple(prob.tspan),
                             solver=_capi.SOLVER_LAMBA_EM if adaptive else _capi.SOLVER_EM,
                             abstol=float(S.kwargs.get("abstol", 1e-2)), reltol=float(S.kwargs.get("reltol", 1e-2)),
                             dtmax=float(S.kwargs.get("dtmax", 0.0)), maxiters=int(S.kwargs.get("maxiters", 100000)),
                             n_kkl=S.n, dt_fixed=dt, pdf=self.d.table(), fp_mode=FP32 if ens.fp32 else FP64,
                             devices=ens.devices, block_size=ens.block_size, blocks_per_sm=ens.blocks_per_sm,
                             refill_threshold=ens.refill_threshold, chunk_points=ens.chunk_points,
                         schedule=ens.schedule)
        ens = S.ensemblealg
        if not isinstance(ens, EnsembleB200):
            raise TypeError("this package only carries the EnsembleB200 ensemble algorithm (no CPU fallback)")
        if S.alg is not None and not isinstance(S.alg, Tsit5):
            raise NotImplementedError("the B200 ODE path implements Tsit5()")
        kw = dict(S.kwargs)
        src = prob.f + "\n" + self.h + "\n" + self.g
        return ModelSpec(source=src, nstate=len(prob.u0), nparam=len(prob.p), nx=len(self.d), nout=self.nout,
                         u0=list(prob.u0), p=list(prob.p), tspan=tuple(prob.tspan), solver=_capi.SOLVER_TSIT5,
                         abstol=float(kw.get("abstol", 1e-6)), reltol=float(kw.get("reltol", 1e-3)),
                         dtmax=float(kw.get("dtmax", 0.0)), maxiters=int(kw.get("maxiters", 100000)),
                         pdf=self.d.table(), fp_mode=FP32 if ens.fp32 else FP64, devices=ens.devices,
                         block_size=ens.block_size, blocks_per_sm=ens.blocks_per_sm,
                         refill_threshold=ens.refill_threshold, chunk_points=ens.chunk_points,
                         schedule=ens.schedule, save_times=tuple(float(t) for t in prob.saveat))

    def handle(self) -> _capi.Handle:
        spec = self.model_spec()
        key = _spec_key(spec)
        if self._handle is None or self._handle_key != key:
            if self._handle is not None:
                self._handle.close()
            self._handle = _capi.Handle(spec)
            self._handle_key = key
        return self._handle

    def close(self):
        if self._handle is not None:
            self._handle.close()
            self._handle = None


def _spec_key(spec) -> str:
    """Cache key of a compiled handle: every structured field of the model, with tabulated density factors
    digested byte for byte (numpy elides the repr of arrays beyond 1000 elements, so two KDE tables with equal
    ends would otherwise share a handle -- and the stale table in HBM)."""
    import dataclasses
    import hashlib
    d = dataclasses.asdict(spec)
    pdf = d.pop("pdf")
    parts = [repr(sorted(d.items()))]
    for f in pdf or ():
        f = tuple(f)
        parts.append(repr(tuple(float(v) for v in f[:5])))
        if len(f) > 5:
            tab = np.ascontiguousarray(f[5], dtype=np.float64)
            parts.append(f"table[{tab.size}]:" + hashlib.sha1(tab.tobytes()).hexdigest())
    return "|".join(parts)


@dataclass
class ExpectationSolution:
    u: Any
    resid: Any
    original: Any

    def __repr__(self):
        u = np.asarray(self.u)
        if u.ndim and u.size > 1:
            return f"ExpectationSolution(u={u.size}-element {type(self.u).__name__})"
        return f"ExpectationSolution(u={self.u})"


@dataclass(frozen=True)
class Koopman:
    sensealg: Any = None


@dataclass
class MonteCarlo:
    """``MonteCarlo(trajectories)``.  The reference draws ``rand(d)`` inside ``prob_func`` for every
    trajectory (expectation.jl:278); here ``seed`` keys a counter-based stream (sample i is a pure
    function of (seed, i)) that the kernel draws in registers (``sampler="device"``, the default:
    nothing crosses PCIe) and ``sample_host`` reproduces bit for bit for a CPU arm.
    ``sampler="host"`` generates the numpy PCG64 block stream on the host instead (block b is a pure
    function of (seed, b)) and ships it; ``samples`` supplies a caller-generated set -- the "same
    host-generated sample set" both arms consume."""
    trajectories: int
    seed: int = 0
    samples: Any = None
    block: int = 1 << 20
    sampler: str = "device"


@dataclass(frozen=True)
class CubatureJLh:
    """h-adaptive Genz-Malik cubature (Integrals.jl's ``CubatureJLh``; error_norm INDIVIDUAL)."""
    pass


@dataclass(frozen=True)
class B200Cubature:
    """The same h-adaptive Genz-Malik algorithm with the region heap inside libb200expect and the rule
    applied on the device (``b200e_koopman_solve``): regions go up, (value, error, split dimension) per
    region come back -- no ``nout x npts`` round trip per refinement round."""
    pass


class BatchIntegralFunction:
    """``BatchIntegralFunction{true}(f!, prototype; max_batch)`` of the reference (expectation.jl:199)."""

    def __init__(self, f, prototype, max_batch):
        self.f = f
        self.integrand_prototype = prototype
        self.max_batch = max_batch

    def __call__(self, dx, x, p=None):
        return self.f(dx, x, p)


def build_integrand(prob: ExpectationProblem, alg: Koopman, mid, p, batch: Optional[int]):
    """``build_integrand(prob, ::Koopman, mid, p, batch::Integer)`` for SystemMap and
    ProcessNoiseSystemMap problems (reference src/expectation.jl:169-200, :209-249).

    Returns ``f!(dx, x, p)``: ``x`` is ``dim x npts`` in Julia, i.e. ``(npts, dim)`` C-order here, and
    ``dx`` is ``(npts, nout)`` -- or ``(npts,)`` for a scalar observable (test/interface.jl:78-98).
    As in the reference, one prototype trajectory at the domain midpoint is solved up front."""
    if batch is None:
        raise NotImplementedError("the unbatched integrand stays on the CPU path of the reference; "
                                  "pass batch=<int> (reference: 'Batching is required' for process noise)")
    h = prob.handle()
    nx, nout = h.spec.nx, h.spec.nout

    def integrand_batch(dx, x, _p=None):
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float64)).reshape(-1, nx)
        out = h.eval_nodes(x, layout=LAYOUT_POINT_MAJOR, weight_by_pdf=True)
        dx[...] = out.reshape(np.shape(dx))
        return None

    mid = np.asarray(mid, dtype=np.float64).reshape(1, nx)
    proto = h.eval_nodes(mid, weight_by_pdf=True)
    prototype = proto.reshape(1) if nout == 1 else proto.reshape(1, nout)
    return BatchIntegralFunction(integrand_batch, prototype, batch)


def solve(prob: ExpectationProblem, expalg, *, maxiters: int = 1000000, batch: Optional[int] = None,
          quadalg=None, ireltol: float = 1e-2, iabstol: float = 1e-2, record_batches: Optional[list] = None):
    """``solve(exprob, Koopman(); maxiters, batch, quadalg, ireltol, iabstol)`` /
    ``solve(exprob, MonteCarlo(n))`` (reference src/expectation.jl:336-354, :268-325).

    ``batch`` selects the batched integrand (``build_integrand(..., batch::Integer)``) and is stored as
    ``max_batch`` of the integrand, as in the reference.  It does NOT cap the size of a refinement round under
    ``CubatureJLh()``: Cubature's ``hcubature_v`` ignores ``max_batch`` (only Cuba's algorithms use it, as
    ``nvec``), so a round holds as many nodes as the region heap asks for -- here too."""
    if isinstance(expalg, MonteCarlo):
        return _solve_montecarlo(prob, expalg)
    if not isinstance(expalg, Koopman):
        raise TypeError(f"unknown expectation algorithm {expalg!r}")
    if quadalg is None:
        quadalg = CubatureJLh()
    lb, ub = extrema(prob.d)
    if isinstance(quadalg, B200Cubature):
        if batch is None:
            raise NotImplementedError("pass batch=<int>: the device path is the batched integrand")
        u, resid, stats = prob.handle().koopman_solve(lb, ub, reltol=ireltol, abstol=iabstol, maxevals=maxiters)
        return ExpectationSolution(u[0] if prob.nout == 1 else u, resid[0] if prob.nout == 1 else resid, stats)
    if not isinstance(quadalg, CubatureJLh):
        raise NotImplementedError("the batched B200 path is driven by CubatureJLh() or B200Cubature()")
    integrand = build_integrand(prob, expalg, (lb + ub) / 2.0, prob.params, batch)
    nout = prob.nout

    def f(x):
        dx = np.empty((len(x), nout))
        integrand(dx, x, prob.params)
        return dx

    res = hcubature_v(f, lb, ub, fdim=nout, reltol=ireltol, abstol=iabstol, maxevals=maxiters,
                      max_batch=None, record=record_batches)
    u = res.u[0] if nout == 1 else res.u
    resid = res.resid[0] if nout == 1 else res.resid
    return ExpectationSolution(u, resid, res)


def _g_powers(g: str, n: int) -> str:
    """``g_powers(sol, p) = [g(sol, p)^i for i in 1:n]`` (reference src/expectation.jl:565-568) for a scalar
    observable given as device source: the user's function is renamed and wrapped."""
    if "void observable_at(" in g:
        raise NotImplementedError("centralmoment of a time-sampled observable: take the moments per save time instead")
    if "void observable(" not in g:
        raise ValueError("g must define `observable(u_end, p, t_end, out)`")
    body = "\n".join(f"    out[{i}] = out[{i - 1}] * v[0];" for i in range(1, n))
    return (g.replace("void observable(", "void b200e_observable_base(")
            + "B200E_FN void observable(const real* u, const real* p, real t_end, real* out) {\n"
              "    real v[1];\n    b200e_observable_base(u, p, t_end, v);\n    out[0] = v[0];\n"
            + body + "\n}\n")


def centralmoment(n: int, exprob: ExpectationProblem, expalg, **kwargs):
    """``centralmoment(n, exprob, expalg; kwargs...)`` (reference src/expectation.jl:554-593): the first ``n``
    central moments of a scalar observable.  One solve of the n-output observable ``[g, g^2, ..., g^n]``
    through the batched path, then the binomial expansion
    ``mu_m = sum_k C(m,k) (-mu)^(m-k) E[X^k]`` on the host, exactly as the reference does."""
    if n < 1:
        raise ValueError("n must be at least 1")
    if exprob.nout != 1:
        raise ValueError("centralmoment needs a scalar observable (nout = 1)")
    new = ExpectationProblem(exprob.S, _g_powers(exprob.g, n), exprob.h,
                             None if isinstance(exprob.S, ProcessNoiseSystemMap) else exprob.d, nout=n)
    try:
        sol = solve(new, expalg, **kwargs)
    finally:
        new.close()
    raw = np.atleast_1d(np.asarray(sol.u, dtype=np.float64))
    mu = raw[0]
    central = np.empty(n)
    for m in range(1, n + 1):
        cm = 0.0
        for k in range(0, m + 1):
            e_xk = 1.0 if k == 0 else raw[k - 1]
            cm += math.comb(m, k) * (-mu) ** (m - k) * e_xk
        central[m - 1] = cm
    return central


def _solve_montecarlo(prob: ExpectationProblem, alg: MonteCarlo):
    h = prob.handle()
    n = int(alg.trajectories)
    nout = prob.nout
    total = np.zeros(nout)
    counters = {"nf": 0, "naccept": 0, "nreject": 0, "nfail": 0}
    if alg.samples is not None:
        s, _, c = h.reduce_samples(alg.samples, weight_by_pdf=False, npts=n)
        total += s
        counters = c
    elif alg.sampler == "device" and all(getattr(f, "device_sampler", True) for f in prob.d.dists):
        s, _, c = h.reduce_generated(alg.seed, n)
        total += s
        counters = c
    elif alg.sampler not in ("host", "device"):
        raise ValueError(f"unknown sampler {alg.sampler!r} (device | host)")
    else:
        done, b = 0, 0
        while done < n:
            m = min(alg.block, n - done)
            x = prob.d.rand_block(alg.seed, b, alg.block)[:m] if m < alg.block else prob.d.rand_block(alg.seed, b, alg.block)
            s, _, c = h.reduce_samples(np.ascontiguousarray(x), weight_by_pdf=False)
            total += s
            for k in counters:
                counters[k] += c[k]
            done += m
            b += 1
    mean = total / n  # mean(sol.u), expectation.jl:293
    sol = ExpectationSolution(mean[0] if nout == 1 else mean, None, None)
    sol.counters = counters
    return sol
