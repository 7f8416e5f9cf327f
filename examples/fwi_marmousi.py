"""examples/nn_fwi/FWI_inversion.jl of the reference in miniature, on libadseis_b200 (needs a B200):
   python examples/fwi_marmousi.py <dir with marmousi2-model-true.mat and marmousi2-model-smooth.mat> [iterations]
Observed data are simulated from the true model (the reference reads them from text files written by its
FWI_forward.jl), the inversion starts from the smooth model with the mean-normalised masked parameterisation of
src/IO.jl:172-197 and runs L-BFGS (src/Optim.jl:135-193).  All shots of the file are used; under torchrun they are
distributed over the GPUs as in compute_loss_and_grads_GPU (src/Utils.jl:300-332)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import adseis_b200 as A  # noqa: E402

d = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/examples/nn_fwi/models"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
f_true, f_smooth = os.path.join(d, "marmousi2-model-true.mat"), os.path.join(d, "marmousi2-model-smooth.mat")
param, vp_true = A.io.load_acoustic_model(f_true, vp_ref=1e3, PropagatorKernel=1)
_, vp0 = A.io.load_acoustic_model(f_smooth)
srcs, rcvs = A.io.load_acoustic_source(f_true), A.io.load_acoustic_receiver(f_true)
ctx = A.default_context()

plans = []
for s, r in zip(srcs, rcvs):                       # one device-resident plan per shot, reused by every iteration
    plan = A.AcousticPlan(param, s.srci, s.srcj, r.rcvi, r.rcvj, ctx=ctx)
    plan.set_srcv(s.srcv)
    plan.set_model(vp_true)
    plan.forward()
    plan.set_obs(plan.rcvv())                      # "observed" data of this shot
    plans.append(plan)

mask = np.ones_like(vp0)
mask[:, :12] = 0                                   # keep the water layer fixed, as the reference's mask does
vp = A.fwi.ConstantOrVariable(vp0, trainable=True, mask=mask).cuda()


def misfit():
    c = vp()
    return sum(A.fwi.acoustic_misfit(p, c) for p in plans)


losses = A.fwi.LBFGS_(misfit, vp.parameters(), max_iter=iters, callback=lambda _, it, L: print("iter %3d  loss %.6e" % (it, L)))
err0 = np.linalg.norm(vp0 - vp_true) / np.linalg.norm(vp_true)
err1 = np.linalg.norm(vp().detach().cpu().numpy() - vp_true) / np.linalg.norm(vp_true)
print("loss %.4e -> %.4e ; model error %.4f -> %.4f" % (losses[0], losses[-1], err0, err1))
