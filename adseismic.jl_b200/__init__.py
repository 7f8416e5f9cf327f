"""adseismic.jl_b200 -- B200-native (sm_100a) drop-in for the FDTD hot path of kailaix/ADSeismic.jl.

Everything numerical lives in libadseis_b200.so (csrc/, C ABI in include/adseis.h); this package is the host-side
mirror of the reference's Julia interface for that path.  Import it as `adseis_b200` (see /adseis_b200.py)."""
from . import _lib
from ._lib import AdseisError, Context, default_context
from .structs import (AcousticPropagatorParams, AcousticReceiver, AcousticSource, ElasticPropagatorParams,
                      ElasticReceiver, ElasticSource)
from .acoustic import (AcousticPlan, AcousticPropagator, AcousticPropagatorSolver,
                       acoustic_forward, acoustic_misfit_grad, acoustic_one_step, acoustic_one_step_grad,
                       compute_PML_Params_)
from . import acoustic as _ac
from . import elastic as _el
from .elastic import (ElasticPlan, ElasticPropagator, ElasticPropagatorSolver, compute_PML_Params, elastic_forward,
                      elastic_misfit_grad)
from .utils import Gauss, Ricker, compute_lame_parameters
from . import io
from . import workloads


def SimulatedObservation_(prop, rcv):
    """SimulatedObservation!(propagator, receiver): src/Core.jl:726-730 (acoustic, rcvv [(NSTEP+1), nrcv]) and
    src/Core.jl:701-712 (elastic, rcvv [nrcv, (NSTEP+1)])."""
    if isinstance(prop, ElasticPropagator):
        rcv.rcvv = elastic_forward(prop.param, prop.src, prop.rho, prop.lam, prop.mu, rcv, False, prop.ctx)[0]
        return rcv.rcvv
    return _ac.SimulatedObservation_(prop, rcv)


def __getattr__(name):
    if name == "fwi":      # needs torch: imported on first use
        import importlib
        return importlib.import_module(__name__ + ".fwi")
    raise AttributeError(name)


def build(force=False, verbose=False):
    """Compile csrc/*.cu into libadseis_b200.so (nvcc, sm_100a, in-tree)."""
    from . import _build
    return _build.build(force=force, verbose=verbose)
