"""gscf-b200: B200-native implementation of the SCF-iteration hot path of molsturm/gscf.

The compute path is libgscf_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/gscf_b200.h).  This package is the Python host-side mirror of gscf's solver /
plugin interface; importing it fails loudly when the shared library has not been built.
"""
from . import lib  # noqa: F401  (raises ImportError when libgscf_b200.so is missing)
from . import observe  # noqa: F401
from .solvers import (  # noqa: F401
    ALL,
    DeviceDFFock,
    Eigensolution,
    Engine,
    ExcDiisStepFailed,
    ExcInnerEigensolverFailed,
    ExcInvalidState,
    ExcMaximumNumberOfIterationsReached,
    FocklikeMatrix_i,
    GenMap,
    OrbitalSpace,
    PlainScf,
    PlainScfKeys,
    PulayDiisScf,
    PulayDiisScfKeys,
    ScfBase,
    ScfBaseKeys,
    ScfState,
    SolverException,
    TruncatedOptDampScf,
    TruncatedOptDampScfKeys,
    solve_device_problem,
)

__version__ = "0.1.0"
