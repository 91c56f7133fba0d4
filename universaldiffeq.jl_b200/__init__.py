"""universaldiffeq.jl_b200 -- B200-native loss-and-gradient engine for UniversalDiffEq-style UDE training.

Host-side mirror of the reference's API for the hot path only: model constructors,
train!, cross-validation drivers, all calling libude_cuda.so through its C ABI (include/ude_cuda.h).
The directory name carries a dot, so import it through the repo-root shim:  ``import ude_b200 as ude``.
"""
from . import problem, synthetic  # noqa: F401
from .problem import Problem  # noqa: F401
from ._capi import Engine, UdeError, device_count, measure_fp64_peak, get_unique_id, load_library, LIB_PATH  # noqa: F401,E402
