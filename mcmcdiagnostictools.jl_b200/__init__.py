"""mcmcdiagnostictools.jl_b200 — B200-native ESS / R-hat hot path of MCMCDiagnosticTools.jl.

The directory name contains a dot, so import it through the root-level loader:

    import mcmcdiag_b200 as mcd
    mcd.ess_rhat(x)            # x: (draws, chains, params...) NumPy array or CUDA tensor
"""
from .api import *  # noqa: F401,F403
from .api import __all__  # noqa: F401
from . import _lib, build, sharding  # noqa: F401

__version__ = "0.1.0"
