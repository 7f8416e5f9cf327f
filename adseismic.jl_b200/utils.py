"""Input builders of the reference (src/Utils.jl): Ricker, Gauss, Lame helpers.  Host-side NumPy (setup only; these
are evaluated once per problem in the reference too, at graph-build time)."""
import numpy as np


def Ricker(epp, a, shift, amp=1.0):
    """src/Utils.jl:188-204: amp*A*(1-t^2/a^2)*exp(-t^2/(2a^2)), t = (1..NSTEP) - shift, A = 2/(sqrt(3a)*pi^(1/4))."""
    NT = epp.NSTEP
    A = 2.0 / (np.sqrt(3.0 * a) * (np.pi ** 0.25))
    wsq = a ** 2
    vec = np.arange(1, NT + 1, dtype=np.float64) - shift
    xsq = vec ** 2
    return amp * A * (1 - xsq / wsq) * np.exp(-xsq / (2 * wsq))


def Gauss(epp, a, shift=None, amp=1.0):
    """src/Utils.jl:206-220"""
    if shift is None:
        shift = 1.2 / a
    t = np.arange(0, epp.NSTEP, dtype=np.float64) * epp.DELTAT
    A = np.pi ** 2 * a ** 2
    return amp * 2.0 * A * np.exp(-A * (t - shift) ** 2)


def compute_lame_parameters(*args):
    """src/Utils.jl:223-243: (NX, NY, vp, vs, rho) -> constant padded arrays; (vp, vs, rho) arrays -> (lam, mu, rho)."""
    if len(args) == 5:
        NX, NY, vp, vs, rho = args
        shape = (NX + 2, NY + 2)
        return (np.full(shape, rho * (vp * vp - 2.0 * vs * vs)), np.full(shape, rho * vs * vs), np.full(shape, float(rho)))
    vp, vs, rho = (np.asarray(x, dtype=np.float64) for x in args)
    return rho * (vp * vp - 2.0 * vs * vs), rho * vs * vs, rho
