"""examples/demo/AcousticWave.jl of the reference, on libadseis_b200 (needs a B200): 150 x 150 cells, 1000 steps,
a Ricker source in the middle, homogeneous 3000 m/s; prints the wavefield energy instead of plotting."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adseis_b200 as A  # noqa: E402

param = A.AcousticPropagatorParams(NX=150, NY=150, NSTEP=1000, DELTAT=1e-4, DELTAX=1.0, DELTAY=1.0, vp_ref=3000.0,
                                   Rcoef=0.001)          # PropagatorKernel=0, the reference's default
rc = A.Ricker(param, 15.0, 100.0, 1e10)
src = A.AcousticSource([param.NX // 2], [param.NY // 2], rc.reshape(-1, 1))
c = 3000.0 * np.ones((param.NX + 2, param.NY + 2))
model = A.AcousticPropagatorSolver(param, src, c)
rcv = A.AcousticReceiver(np.arange(10, 140), np.full(130, 20))
A.SimulatedObservation_(model, rcv)
u = model.u                                               # (NSTEP+1, NX+2, NY+2), copied from the device on demand
print("traces", rcv.rcvv.shape, "max |trace| %.3e" % np.abs(rcv.rcvv).max())
for s in (250, 500, 1000):
    print("step %4d  sum u^2 = %.6e" % (s, float((u[s] ** 2).sum())))
