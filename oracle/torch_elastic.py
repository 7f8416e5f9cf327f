"""PyTorch-autograd restatement of the reference's ELASTIC graph -- TEST INFRASTRUCTURE ONLY.

The reference has no hand-written elastic adjoint: its gradient is tf.gradients through the op composition of
src/Core.jl:31-228 (variant "S") / src/MPIElastic.jl:374-645 (variant "M", one block).  This module rebuilds that
composition literally -- every `x[idx]` is a gather (GatherOps.h:1-5), every `scatter_add_op` copies and adds
(ScatterAddOps.h:1-7), every `makevector` zero-fills and sets (ScatterNdOps.h:1-6, .cpp:92) -- on torch fp64 CPU
tensors and lets torch.autograd play the role of tf.gradients.  It is used (here, where torch is available) to
pin oracle/oracle.c's hand-derived reverse sweep and to generate tests/golden/elastic_*.npz.
"""
import numpy as np
import torch


def _getid(a, b, W):
    """Core.jl:669-677 / MPIElastic.jl:146-154; returns 0-based flat indices (1-based ranges a, b inclusive)."""
    ii = np.arange(a[0], a[1] + 1)[:, None]
    jj = np.arange(b[0], b[1] + 1)[None, :]
    return torch.as_tensor(((ii - 1) * W + (jj - 1)).reshape(-1), dtype=torch.int64)


def _scatter_add(ipt, ii, upd):
    return ipt.index_add(0, ii, upd)


def _makevector(m_len, ii, o):
    return torch.zeros(m_len, dtype=torch.float64).index_copy(0, ii, o)


def _bx(a, nrep):   # adbroadcast(a, b, 1): coefficient varies with the slow (x) index
    return a.repeat_interleave(nrep)


def _by(a, nrep):   # adbroadcast(a, b, 2): coefficient varies with the fast (y) index
    return a.repeat(nrep)


def elastic_loss(variant, NX, NY, NSTEP, dt, dx, dy, ax, bx, ay, by, rho, lam, mu, srci, srcj, srctype, srcv, rcvi,
                 rcvj, rcvtype, obs):
    """rho/lam/mu: flat torch tensors of the padded H*W arrays; srcv torch [>=NSTEP, nsrc].
    Returns (loss, rcvv[nrcv, NSTEP+1])."""
    t = lambda x: torch.as_tensor(np.asarray(x), dtype=torch.float64)
    ax, bx, ay, by = t(ax).reshape(2, NX), t(bx).reshape(2, NX), t(ay).reshape(2, NY), t(by).reshape(2, NY)
    if variant == 0:
        H, W = NX + 2, NY + 2
        soff = -1
    else:
        H, W = NX + 4, NY + 4
        soff = 1
    HW = H * W
    z = lambda: torch.zeros(HW, dtype=torch.float64)
    vx, vy, sxx, syy, sxy = z(), z(), z(), z(), z()
    mem = [z() for _ in range(8)]
    fields_hist = [[vx, vy, sxx, syy, sxy]]

    if variant == 0:
        # fw1 (Core.jl:100-107)
        f1 = dict(i_1j=_getid((1, NX - 1), (3, NY + 1), W), ij=_getid((2, NX), (3, NY + 1), W),
                  i1j=_getid((3, NX + 1), (3, NY + 1), W), i2j=_getid((4, NX + 2), (3, NY + 1), W),
                  ij_2=_getid((2, NX), (1, NY - 1), W), ij_1=_getid((2, NX), (2, NY), W),
                  ij1=_getid((2, NX), (4, NY + 2), W))
        f2 = dict(i_2j=_getid((1, NX - 1), (2, NY), W), i_1j=_getid((2, NX), (2, NY), W),
                  ij=_getid((3, NX + 1), (2, NY), W), i1j=_getid((4, NX + 2), (2, NY), W),
                  ij_1=_getid((3, NX + 1), (1, NY - 1), W), ij1=_getid((3, NX + 1), (3, NY + 1), W),
                  ij2=_getid((3, NX + 1), (4, NY + 2), W))
        f3 = dict(i_2j=_getid((1, NX - 1), (3, NY + 1), W), i_1j=_getid((2, NX), (3, NY + 1), W),
                  ij=_getid((3, NX + 1), (3, NY + 1), W), i1j=_getid((4, NX + 2), (3, NY + 1), W),
                  ij_2=_getid((3, NX + 1), (1, NY - 1), W), ij_1=_getid((3, NX + 1), (2, NY), W),
                  ij1=_getid((3, NX + 1), (4, NY + 2), W))
        f4 = dict(i_1j=_getid((1, NX - 1), (2, NY), W), ij=_getid((2, NX), (2, NY), W),
                  i1j=_getid((3, NX + 1), (2, NY), W), i2j=_getid((4, NX + 2), (2, NY), W),
                  i1j1=_getid((3, NX + 1), (3, NY + 1), W), ij_1=_getid((2, NX), (1, NY - 1), W),
                  ij1=_getid((2, NX), (3, NY + 1), W), ij2=_getid((2, NX), (4, NY + 2), W))
        nxr, nyr = NX - 1, NY - 1  # region extents
        cx1 = (bx[1, :-1], ax[1, :-1]); cy1 = (by[0, 1:], ay[0, 1:])
        cx2 = (bx[0, 1:], ax[0, 1:]);   cy2 = (by[1, :-1], ay[1, :-1])
        cx3 = (bx[0, 1:], ax[0, 1:]);   cy3 = (by[0, 1:], ay[0, 1:])
        cx4 = (bx[1, :-1], ax[1, :-1]); cy4 = (by[1, :-1], ay[1, :-1])
    else:
        n1, n2 = NX, NY  # MPIElastic.jl:175-189 with a (possibly non-square) single block
        k1x, k2x, k_1x, k_2x, kkx = (4, n1 + 3), (5, n1 + 4), (2, n1 + 1), (1, n1), (3, n1 + 2)
        k1y, k2y, k_1y, k_2y, kky = (4, n2 + 3), (5, n2 + 4), (2, n2 + 1), (1, n2), (3, n2 + 2)
        com = dict(i_1j=_getid(k_1x, kky, W), i_2j=_getid(k_2x, kky, W), i1j=_getid(k1x, kky, W),
                   i2j=_getid(k2x, kky, W), ij_1=_getid(kkx, k_1y, W), ij_2=_getid(kkx, k_2y, W),
                   ij1=_getid(kkx, k1y, W), ij2=_getid(kkx, k2y, W), ij=_getid(kkx, kky, W),
                   i1j1=_getid(k1x, k1y, W))
        f1 = f2 = f3 = f4 = com
        nxr, nyr = NX, NY
        cx1 = (bx[1], ax[1]); cy1 = (by[0], ay[0])
        cx2 = (bx[0], ax[0]); cy2 = (by[1], ay[1])
        cx3 = (bx[0], ax[0]); cy3 = (by[0], ay[0])
        cx4 = (bx[1], ax[1]); cy4 = (by[1], ay[1])

    avg = variant == 0
    srcidx = torch.as_tensor([(int(i) + soff) * W + (int(j) + soff) for i, j in zip(srci, srcj)], dtype=torch.int64)
    rcvidx = [(int(i) + soff) * W + (int(j) + soff) for i, j in zip(rcvi, rcvj)]

    for s in range(1, NSTEP + 1):
        # ---- fw1
        d = f1; ij = d["ij"]
        l_ = 0.5 * (lam[d["i1j"]] + lam[ij]) if avg else lam[ij]
        m_ = 0.5 * (mu[d["i1j"]] + mu[ij]) if avg else mu[ij]
        lm = l_ + 2 * m_
        dvx_dx = (27 * vx[d["i1j"]] - 27 * vx[ij] - vx[d["i2j"]] + vx[d["i_1j"]]) / (24 * dx)
        dvy_dy = (27 * vy[ij] - 27 * vy[d["ij_1"]] - vy[d["ij1"]] + vy[d["ij_2"]]) / (24 * dy)
        mem[0] = _makevector(HW, ij, _bx(cx1[0], nyr) * mem[0][ij] + _bx(cx1[1], nyr) * dvx_dx)
        mem[1] = _makevector(HW, ij, _by(cy1[0], nxr) * mem[1][ij] + _by(cy1[1], nxr) * dvy_dy)
        dvx_dx = dvx_dx + mem[0][ij]
        dvy_dy = dvy_dy + mem[1][ij]
        sxx = _scatter_add(sxx, ij, (lm * dvx_dx + l_ * dvy_dy) * dt)
        syy = _scatter_add(syy, ij, (lm * dvy_dy + l_ * dvx_dx) * dt)
        # ---- fw2
        d = f2; ij = d["ij"]
        m_ = 0.5 * (mu[ij] + mu[d["ij1"]]) if avg else mu[ij]
        dvy_dx = (27 * vy[ij] - 27 * vy[d["i_1j"]] - vy[d["i1j"]] + vy[d["i_2j"]]) / (24 * dx)
        dvx_dy = (27 * vx[d["ij1"]] - 27 * vx[ij] - vx[d["ij2"]] + vx[d["ij_1"]]) / (24 * dy)
        mem[2] = _makevector(HW, ij, _bx(cx2[0], nyr) * mem[2][ij] + _bx(cx2[1], nyr) * dvy_dx)
        mem[3] = _makevector(HW, ij, _by(cy2[0], nxr) * mem[3][ij] + _by(cy2[1], nxr) * dvx_dy)
        dvy_dx = dvy_dx + mem[2][ij]
        dvx_dy = dvx_dy + mem[3][ij]
        sxy = _scatter_add(sxy, ij, m_ * (dvy_dx + dvx_dy) * dt)
        # ---- fw3
        d = f3; ij = d["ij"]
        dsxx_dx = (27 * sxx[ij] - 27 * sxx[d["i_1j"]] - sxx[d["i1j"]] + sxx[d["i_2j"]]) / (24 * dx)
        dsxy_dy = (27 * sxy[ij] - 27 * sxy[d["ij_1"]] - sxy[d["ij1"]] + sxy[d["ij_2"]]) / (24 * dy)
        mem[4] = _makevector(HW, ij, _bx(cx3[0], nyr) * mem[4][ij] + _bx(cx3[1], nyr) * dsxx_dx)
        mem[5] = _makevector(HW, ij, _by(cy3[0], nxr) * mem[5][ij] + _by(cy3[1], nxr) * dsxy_dy)
        dsxx_dx = dsxx_dx + mem[4][ij]
        dsxy_dy = dsxy_dy + mem[5][ij]
        vx = _scatter_add(vx, ij, (dsxx_dx + dsxy_dy) * dt / rho[ij])
        # ---- fw4
        d = f4; ij = d["ij"]
        r_ = 0.25 * (rho[ij] + rho[d["i1j"]] + rho[d["i1j1"]] + rho[d["ij1"]]) if avg else rho[ij]
        dsxy_dx = (27 * sxy[d["i1j"]] - 27 * sxy[ij] - sxy[d["i2j"]] + sxy[d["i_1j"]]) / (24 * dx)
        dsyy_dy = (27 * syy[d["ij1"]] - 27 * syy[ij] - syy[d["ij2"]] + syy[d["ij_1"]]) / (24 * dy)
        mem[6] = _makevector(HW, ij, _bx(cx4[0], nyr) * mem[6][ij] + _bx(cx4[1], nyr) * dsxy_dx)
        mem[7] = _makevector(HW, ij, _by(cy4[0], nxr) * mem[7][ij] + _by(cy4[1], nxr) * dsyy_dy)
        dsxy_dx = dsxy_dx + mem[6][ij]
        dsyy_dy = dsyy_dy + mem[7][ij]
        vy = _scatter_add(vy, ij, (dsxy_dx + dsyy_dy) * dt / r_)
        # ---- add_source (AddSource.cpp:33-87): sequential +=
        fl = [vx, vy, sxx, syy, sxy]
        for k in range(len(srci)):
            ty = int(srctype[k])
            if 0 <= ty <= 4:
                fl[ty] = fl[ty].index_add(0, srcidx[k:k + 1], srcv[s - 1, k:k + 1])
        vx, vy, sxx, syy, sxy = fl
        fields_hist.append([vx, vy, sxx, syy, sxy])

    rows = []
    for r in range(len(rcvi)):
        ty = int(rcvtype[r])
        rows.append(torch.stack([fields_hist[s][ty][rcvidx[r]] for s in range(NSTEP + 1)]))
    rcvv = torch.stack(rows) if rows else torch.zeros((0, NSTEP + 1), dtype=torch.float64)
    loss = ((rcvv - torch.as_tensor(np.asarray(obs), dtype=torch.float64)) ** 2).sum()
    return loss, rcvv


def elastic_misfit_grad(variant, NX, NY, NSTEP, dt, dx, dy, ax, bx, ay, by, rho, lam, mu, srci, srcj, srctype, srcv,
                        rcvi, rcvj, rcvtype, obs):
    rho_t = torch.tensor(np.asarray(rho, dtype=np.float64).reshape(-1), requires_grad=True)
    lam_t = torch.tensor(np.asarray(lam, dtype=np.float64).reshape(-1), requires_grad=True)
    mu_t = torch.tensor(np.asarray(mu, dtype=np.float64).reshape(-1), requires_grad=True)
    srcv_t = torch.tensor(np.asarray(srcv, dtype=np.float64)[:NSTEP], requires_grad=True)
    loss, rcvv = elastic_loss(variant, NX, NY, NSTEP, dt, dx, dy, ax, bx, ay, by, rho_t, lam_t, mu_t, srci, srcj,
                              srctype, srcv_t, rcvi, rcvj, rcvtype, obs)
    loss.backward()
    H, W = (NX + 2, NY + 2) if variant == 0 else (NX + 4, NY + 4)
    g = lambda x: (x.grad.numpy().copy() if x.grad is not None else np.zeros(x.shape))
    return dict(loss=float(loss.detach()), rcvv=rcvv.detach().numpy(), grad_rho=g(rho_t).reshape(H, W),
                grad_lam=g(lam_t).reshape(H, W), grad_mu=g(mu_t).reshape(H, W), grad_srcv=g(srcv_t))
