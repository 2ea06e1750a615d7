"""Drop-in import path: ``from ARBTools.ARBInterp import tricubic, quadcubic`` resolves to the
B200-native classes (the reference ships this module as src/ARBInterp/ARBInterp.py and, in its
wheel, as ARBTools/ARBInterp.py)."""
from arbinterp_b200.interp import tricubic, quadcubic, __version__  # noqa: F401
