"""GPU tests of the layers either side of the hot path: the reference's Marmousi fixture through the
MAT loader conventions (golden file from the reference's op bodies), the torch.autograd misfit wrappers (chain rule
through the parameterisation == library gradient), and an L-BFGS inversion loop that actually reduces the misfit."""
import numpy as np
import pytest

from conftest import golden, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _marmousi(A):
    G = golden("acoustic_marmousi2_shot3.npz")
    p = A.AcousticPropagatorParams(PropagatorKernel=1, NX=int(G["NX"]), NY=int(G["NY"]), NSTEP=int(G["NSTEP"]), DELTAX=float(G["dx"]),
                                   DELTAY=float(G["dy"]), DELTAT=float(G["dt"]), NPOINTS_PML=int(G["npml"]),
                                   vp_ref=float(G["vp_ref"]))
    return G, p


def test_marmousi_golden(A, ctx):
    G, p = _marmousi(A)
    src, rcv = A.AcousticSource(G["srci"], G["srcj"], G["srcv"]), A.AcousticReceiver(G["rcvi"], G["rcvj"])
    R = A.acoustic_misfit_grad(p, src, G["c"], rcv, G["obs"], ctx=ctx)
    assert np.array_equal(R["rcvv"], G["rcvv"])                 # bit-identical traces
    assert abs(R["loss"] - float(G["loss"])) / float(G["loss"]) < 1e-13
    assert relerr(R["grad_c"], G["grad_c"]) < TOL and relerr(R["grad_srcv"], G["grad_srcv"]) < TOL


@pytest.mark.parametrize("device", ["cpu", "cuda"])
def test_acoustic_autograd_chain_rule(A, ctx, device):
    import torch
    G, p = _marmousi(A)
    plan = A.AcousticPlan(p, G["srci"], G["srcj"], G["rcvi"], G["rcvj"], ctx=ctx)
    plan.set_srcv(G["srcv"]); plan.set_obs(G["obs"])
    mask = np.zeros_like(G["c"]); mask[:, 20:] = 1
    vp = A.fwi.ConstantOrVariable(G["c"], trainable=True, mask=mask).to(device)
    srcv = torch.tensor(G["srcv"], device=device, requires_grad=True)
    loss = A.fwi.acoustic_misfit(plan, vp(), srcv)
    assert abs(float(loss.detach()) - float(G["loss"])) / float(G["loss"]) < 1e-13
    (3.0 * loss).backward()
    # d/dx_ [ mask x_ + x0 (1 - mask) ] * mean  ->  grad_c * mask * mean
    want = 3.0 * G["grad_c"] * mask * G["c"].mean()
    assert relerr(vp.x_.grad.cpu().numpy(), want) < TOL
    assert relerr(srcv.grad.cpu().numpy(), 3.0 * G["grad_srcv"]) < TOL
    plan.close()


def test_lbfgs_reduces_marmousi_misfit(A, ctx):
    """examples/nn_fwi/FWI_inversion.jl in miniature: start from the smooth model, fit the true model's traces."""
    import torch
    G, p = _marmousi(A)
    plan = A.AcousticPlan(p, G["srci"], G["srcj"], G["rcvi"], G["rcvj"], ctx=ctx)
    plan.set_srcv(G["srcv"]); plan.set_obs(G["obs"])
    vp = A.fwi.ConstantOrVariable(G["c"], trainable=True).to("cuda")
    seen = []
    losses = A.fwi.LBFGS_(lambda: A.fwi.acoustic_misfit(plan, vp()), vp.parameters(), max_iter=8,
                          callback=lambda params, it, L: seen.append((it, L)))
    assert len(seen) == len(losses) - 1 and seen[-1][1] == losses[-1]
    assert abs(losses[0] - float(G["loss"])) / float(G["loss"]) < 1e-13 and losses[-1] < 0.5 * losses[0]
    assert all(b <= a * (1 + 1e-12) for a, b in zip(losses, losses[1:]))
    plan.close()


def test_elastic_autograd(A, ctx):
    import torch
    G = golden("elastic_S.npz")
    p = A.ElasticPropagatorParams(NX=int(G["NX"]), NY=int(G["NY"]), NSTEP=int(G["NSTEP"]), DELTAX=float(G["dx"]),
                                  DELTAY=float(G["dy"]), DELTAT=float(G["dt"]), NPOINTS_PML=int(G["npml"]),
                                  vp_ref=float(G["vp_ref"]), ALPHA_MAX_PML=float(G["alpha_max"]), variant=0)
    plan = A.ElasticPlan(p, G["srci"], G["srcj"], G["srctype"], G["rcvi"], G["rcvj"], G["rcvtype"], ctx=ctx)
    plan.set_srcv(G["srcv"]); plan.set_obs(G["obs"])
    rho, lam, mu = (torch.tensor(G[k], requires_grad=True) for k in ("rho", "lam", "mu"))
    srcv = torch.tensor(G["srcv"], requires_grad=True)
    loss = A.fwi.elastic_misfit(plan, rho, lam, mu, srcv)
    loss.backward()
    assert abs(float(loss.detach()) - float(G["loss"])) / float(G["loss"]) < 1e-12
    for t, k in ((rho, "grad_rho"), (lam, "grad_lam"), (mu, "grad_mu"), (srcv, "grad_srcv")):
        assert relerr(t.grad.numpy(), G[k]) < TOL, k
    # source-time-function inversion only (rupture-style): materials constant -> no forward history, same grad_srcv
    s2 = torch.tensor(G["srcv"], requires_grad=True)
    A.fwi.elastic_misfit(plan, rho.detach(), lam.detach(), mu.detach(), s2).backward()
    assert relerr(s2.grad.numpy(), G["grad_srcv"]) < TOL
    plan.close()
