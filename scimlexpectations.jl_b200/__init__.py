"""B200-native drop-in for the batched expectation integrand of SciMLExpectations.jl.

The directory name contains a dot, so import it through the alias module at the repo root:
``import scimlexpectations_jl_b200 as sme``.
"""
from . import _capi, models  # noqa: F401
from ._capi import (B200Error, DeviceArray, Handle, ModelSpec, PinnedArray, compile_check, device_count,  # noqa: F401
                    load_library, measure_fma_peak)

__version__ = "0.1.0"
