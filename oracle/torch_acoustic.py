"""PyTorch-autograd restatement of the reference's ACOUSTIC graph for PropagatorKernel=0 -- TEST INFRASTRUCTURE ONLY.

PropagatorKernel=0 has no custom op and no hand-written adjoint in the reference: `one_step` (src/Core.jl:528-549) is
a composition of gathers `x[IJ]` and `scatter_nd_ops`, looped by AcousticPropagatorSolver (src/Core.jl:562-620), and
its gradient is tf.gradients through that graph.  This module rebuilds the composition literally on torch fp64 CPU
tensors (index sets as in compute_PML_Params!, src/Core.jl:656-666) and lets torch.autograd stand in for
tf.gradients.  It pins oracle/oracle.c's hand-derived `orc_acoustic_*_k0` functions and generates
tests/golden/acoustic_kernel0.npz.  kernel=2 is `acoustic_one_step_customop_ref` (src/Core.jl:504-525), the op-free
twin of the custom op (phi, psi from the OLD wavefield), used to cross-check this file against the C++ op bodies."""
import numpy as np
import torch


def _ids(NX, NY):
    W = NY + 2
    ii = np.arange(2, NX + 2)[:, None]           # 1-based interior rows 2..NX+1
    jj = np.arange(2, NY + 2)[None, :]
    f = lambda a, b: torch.as_tensor(((a - 1) * W + (b - 1) + 0 * (ii + jj)).reshape(-1), dtype=torch.int64)
    return dict(IJ=f(ii, jj), IpJ=f(ii + 1, jj), InJ=f(ii - 1, jj), IJp=f(ii, jj + 1), IJn=f(ii, jj - 1))


def one_step(kernel, ix, N, dt, hx, hy, w, wold, phi, psi, sig, tau, c):
    IJ, IpJ, InJ, IJp, IJn = ix["IJ"], ix["IpJ"], ix["InJ"], ix["IJp"], ix["IJn"]
    scat = lambda v: torch.zeros(N, dtype=torch.float64).index_copy(0, IJ, v)
    u = (2 - sig[IJ] * tau[IJ] * dt ** 2 - 2 * dt ** 2 / hx ** 2 * c[IJ] - 2 * dt ** 2 / hy ** 2 * c[IJ]) * w[IJ] + \
        c[IJ] * (dt / hx) ** 2 * (w[IpJ] + w[InJ]) + \
        c[IJ] * (dt / hy) ** 2 * (w[IJp] + w[IJn]) + \
        (dt ** 2 / (2 * hx)) * (phi[IpJ] - phi[InJ]) + \
        (dt ** 2 / (2 * hy)) * (psi[IJp] - psi[IJn]) - \
        (1 - (sig[IJ] + tau[IJ]) * dt / 2) * wold[IJ]
    u = u / (1 + (sig[IJ] + tau[IJ]) / 2 * dt)
    u = scat(u)
    d = u if kernel == 0 else w
    phin = (1. - dt * sig[IJ]) * phi[IJ] + dt * c[IJ] * (tau[IJ] - sig[IJ]) / (2 * hx) * (d[IpJ] - d[InJ])
    psin = (1. - dt * tau[IJ]) * psi[IJ] + dt * c[IJ] * (sig[IJ] - tau[IJ]) / (2 * hy) * (d[IJp] - d[IJn])
    return u, scat(phin), scat(psin)


def acoustic_loss(kernel, NX, NY, NSTEP, dt, hx, hy, sigma, tau, c, srci, srcj, srcv, rcvi, rcvj, obs):
    """c: flat torch tensor of the padded velocity (squared here, Core.jl:564); srcv torch [>=NSTEP, nsrc].
    Returns (loss, rcvv[NSTEP+1, nrcv]) as torch tensors."""
    W, N = NY + 2, (NX + 2) * (NY + 2)
    t = lambda x: torch.as_tensor(np.asarray(x, dtype=np.float64).reshape(-1))
    sig, ta = t(sigma), t(tau)
    ix = _ids(NX, NY)
    c2 = c ** 2
    sidx = torch.as_tensor((np.asarray(srci) - 1) * W + (np.asarray(srcj) - 1), dtype=torch.int64)
    ridx = torch.as_tensor((np.asarray(rcvi) - 1) * W + (np.asarray(rcvj) - 1), dtype=torch.int64)
    z = torch.zeros(N, dtype=torch.float64)
    us, phi, psi = [z, z], z, z
    for s in range(2, NSTEP + 1):
        u, phi, psi = one_step(kernel, ix, N, dt, hx, hy, us[s - 1], us[s - 2], phi, psi, sig, ta, c2)
        u = u.index_add(0, sidx, srcv[s - 1] * dt ** 2)          # scatter_add_op, Core.jl:600-601
        us.append(u)
    rcvv = torch.stack([u[ridx] for u in us])                     # Core.jl:726-730
    loss = ((rcvv - torch.as_tensor(np.asarray(obs, dtype=np.float64))) ** 2).sum()
    return loss, rcvv
