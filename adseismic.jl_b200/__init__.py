"""adseismic.jl_b200 -- B200-native (sm_100a) drop-in for the FDTD hot path of kailaix/ADSeismic.jl.

Everything numerical lives in libadseis_b200.so (csrc/, C ABI in include/adseis.h); this package is the host-side
mirror of the reference's Julia interface for that path.  Import it as `adseis_b200` (see /adseis_b200.py)."""
from . import _lib
from ._lib import AdseisError, Context, default_context
from .structs import (AcousticPropagatorParams, AcousticReceiver, AcousticSource, ElasticPropagatorParams,
                      ElasticReceiver, ElasticSource)
from .acoustic import (AcousticPlan, AcousticPropagator, AcousticPropagatorSolver, SimulatedObservation_,
                       acoustic_forward, acoustic_misfit_grad, acoustic_one_step, acoustic_one_step_grad,
                       compute_PML_Params_)
from .utils import Gauss, Ricker, compute_lame_parameters


def build(force=False, verbose=False):
    """Compile csrc/*.cu into libadseis_b200.so (nvcc, sm_100a, in-tree)."""
    from . import _build
    return _build.build(force=force, verbose=verbose)
