"""gmxapi_b200: B200-native (sm_100a) short-range nonbonded path behind the nblib / nbnxm interface.

Only the hot path of kassonlab/gmxapi's nbnxm module is implemented here. The package is a
thin host-side mirror of the reference's operator interface on top of the C-ABI library libb200nb.so; all
compute runs in hand-written CUDA kernels, and importing the compute classes without the library or without
a GPU fails loudly -- there is no CPU fallback.
"""
from . import systems  # noqa: F401
from .lib import B200NBError, NbnxmGpu, load_library, library_path  # noqa: F401
from .nblib import (CoulombType, ForceCalculator, LjPme, NBKernelOptions, SimulationState, VdwModifier)  # noqa: F401

__all__ = ["systems", "B200NBError", "NbnxmGpu", "load_library", "library_path", "CoulombType", "ForceCalculator",
           "NBKernelOptions", "SimulationState", "VdwModifier", "LjPme"]
