"""B200-native drop-in for the batched expectation integrand of SciMLExpectations.jl.

The directory name contains a dot, so import it through the alias module at the repo root:
``import scimlexpectations_jl_b200 as sme``.
"""
from . import _capi, api, cubature, models  # noqa: F401
from ._capi import (B200Error, DeviceArray, Handle, ModelSpec, PinnedArray, compile_check, device_count, sample_host,  # noqa: F401
                    load_library, measure_fma_peak)
from .api import (EM, B200Cubature, BatchIntegralFunction, Beta, Chi, Chisq, CubatureJLh, EnsembleB200, Erlang,  # noqa: F401
                  ExpectationProblem, ExpectationSolution, Exponential, Gamma, GenericDistribution, Koopman, LambaEM,
                  LogNormal, MonteCarlo, Normal, ODEProblem, Pareto, ProcessNoiseSystemMap, Rayleigh, SDEProblem, SystemMap,
                  Truncated, TruncatedLogPoly, Tsit5, Uniform, UnivariateKDE, build_integrand, centralmoment, extrema,
                  maximum, minimum, pdf, rand, solve, truncated)
from .cubature import hcubature_v  # noqa: F401

__version__ = "0.1.0"
