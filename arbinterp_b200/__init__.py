"""arbinterp_b200 -- B200-native (sm_100a) implementation of ARBInterp's interpolation hot path.

Public surface = the reference's: ``tricubic`` / ``quadcubic`` with ``.Query(points)``.
The CUDA library is loaded lazily (on first construction); importing the package does not need a GPU.
"""
from .interp import tricubic, quadcubic, __version__  # noqa: F401

__all__ = ["tricubic", "quadcubic", "__version__"]
