"""examples/demo/ElasticWave.jl of the reference, on libadseis_b200 (needs a B200): 150 x 150 cells, 500 steps, a
velocity source (type 0) in the middle, homogeneous vp 3000 / vs 1732 / rho 2800; prints field energies instead of
saving an animation, then the misfit gradient w.r.t. (rho, lambda, mu) for a perturbed model."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adseis_b200 as A  # noqa: E402

param = A.ElasticPropagatorParams(NX=150, NY=150, NSTEP=500, DELTAT=1e-4, DELTAX=1.0, DELTAY=1.0, vp_ref=3300.0)
source = A.Ricker(param, 15.0, 100.0, 1e6)
src = A.ElasticSource([param.NX // 2], [param.NY // 2], [0], source.reshape(-1, 1))      # 0: velocity, 1: stress
lam, mu, rho = A.compute_lame_parameters(param.NX, param.NY, 3000.0, 3000.0 / 1.732, 2800.0)
model = A.ElasticPropagatorSolver(param, src, rho, lam, mu)
vx = model.vx                                             # (NSTEP+1, NX+2, NY+2)
for s in (125, 250, 500):
    print("step %4d  sum vx^2 = %.6e" % (s, float((vx[s] ** 2).sum())))

rcv = A.ElasticReceiver(np.arange(20, 130), np.full(110, 30), np.zeros(110, dtype=np.int64))
A.SimulatedObservation_(model, rcv)                       # rcv.rcvv: (nrcv, NSTEP+1)
out = A.elastic_misfit_grad(param, src, rho, 1.05 * lam, mu, rcv, rcv.rcvv)               # examples/demo/ElasticWave_gradtest.jl
print("loss %.6e  |grad_lambda| %.3e  |grad_mu| %.3e  |grad_rho| %.3e" %
      (out["loss"], np.abs(out["grad_lambda"]).max(), np.abs(out["grad_mu"]).max(), np.abs(out["grad_rho"]).max()))
