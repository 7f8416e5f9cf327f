"""NumPy restatement of the reference's BLOCK-DECOMPOSED acoustic solver for PropagatorKernel=0 -- TEST
INFRASTRUCTURE ONLY.  src/MPIAcoustic.jl:212-246 (`one_step`: halo exchanges of w, wold, phi, psi, the update of the
n x n block, a halo exchange of the NEW wavefield, then phi / psi from it) inside the loop of
src/MPIAcoustic.jl:334-404 (source injection after the step), with the M x N ranks held in one process and
`mpi_halo_exchange` (ADCME: edge neighbours, fill value 0 at physical boundaries) emulated by array copies.  Used to
check the reference's own distributed invariant -- decomposed == undecomposed -- for the scheme the oracle restates
on the global grid (orc_acoustic_forward_k0 with mpi_convention)."""
import numpy as np


def _halo(blocks, M, N, n):
    """(n+2) x (n+2) arrays: each block plus one row / column of its edge neighbours (0 outside the domain)."""
    out = [[np.zeros((n + 2, n + 2)) for _ in range(N)] for _ in range(M)]
    for I in range(M):
        for J in range(N):
            p = out[I][J]
            p[1:-1, 1:-1] = blocks[I][J]
            if I > 0: p[0, 1:-1] = blocks[I - 1][J][-1, :]
            if I < M - 1: p[-1, 1:-1] = blocks[I + 1][J][0, :]
            if J > 0: p[1:-1, 0] = blocks[I][J - 1][:, -1]
            if J < N - 1: p[1:-1, -1] = blocks[I][J + 1][:, 0]
    return out


def mpi_acoustic_forward_k0(NX, NY, n, NSTEP, dt, hx, hy, sigma, tau, c2, srci, srcj, srcv):
    """sigma, tau: global padded (NX+2) x (NY+2) profiles; c2: global NX x NY (c squared, MPIAcoustic.jl:336);
    srci, srcj: 1-based into the unpadded global grid.  -> u[(NSTEP+1), NX, NY]"""
    M, N = NX // n, NY // n
    assert M * n == NX and N * n == NY
    sigma, tau = np.asarray(sigma).reshape(NX + 2, NY + 2), np.asarray(tau).reshape(NX + 2, NY + 2)
    blk = lambda a, I, J: a[I * n:(I + 1) * n, J * n:(J + 1) * n]
    S = [[sigma[I * n + 1:I * n + n + 1, J * n + 1:J * n + n + 1] for J in range(N)] for I in range(M)]   # sigma[IJ]
    T = [[tau[I * n + 1:I * n + n + 1, J * n + 1:J * n + n + 1] for J in range(N)] for I in range(M)]
    C = [[blk(c2, I, J) for J in range(N)] for I in range(M)]
    zeros = lambda: [[np.zeros((n, n)) for _ in range(N)] for _ in range(M)]
    u_hist = np.zeros((NSTEP + 1, NX, NY))
    w, wold, phi, psi = zeros(), zeros(), zeros(), zeros()
    for s in range(2, NSTEP + 1):
        W, WO, PH, PS = _halo(w, M, N, n), _halo(wold, M, N, n), _halo(phi, M, N, n), _halo(psi, M, N, n)
        unew = zeros()
        for I in range(M):
            for J in range(N):
                sg, ta, c = S[I][J], T[I][J], C[I][J]
                p, po_, ph, ps = W[I][J], WO[I][J], PH[I][J], PS[I][J]
                IJ = (slice(1, -1), slice(1, -1))
                IpJ, InJ = (slice(2, None), slice(1, -1)), (slice(0, -2), slice(1, -1))
                IJp, IJn = (slice(1, -1), slice(2, None)), (slice(1, -1), slice(0, -2))
                u = (2 - sg * ta * dt ** 2 - 2 * dt ** 2 / hx ** 2 * c - 2 * dt ** 2 / hy ** 2 * c) * p[IJ] + \
                    c * (dt / hx) ** 2 * (p[IpJ] + p[InJ]) + \
                    c * (dt / hy) ** 2 * (p[IJp] + p[IJn]) + \
                    (dt ** 2 / (2 * hx)) * (ph[IpJ] - ph[InJ]) + \
                    (dt ** 2 / (2 * hy)) * (ps[IJp] - ps[IJn]) - \
                    (1 - (sg + ta) * dt / 2) * po_[IJ]
                unew[I][J] = u / (1 + (sg + ta) / 2 * dt)
        U = _halo(unew, M, N, n)                      # halo exchange of the new wavefield (tag 5i+4), BEFORE injection
        nphi, npsi = zeros(), zeros()
        for I in range(M):
            for J in range(N):
                sg, ta, c, uu = S[I][J], T[I][J], C[I][J], U[I][J]
                nphi[I][J] = (1. - dt * sg) * phi[I][J] + dt * c * (ta - sg) / (2 * hx) * (uu[2:, 1:-1] - uu[:-2, 1:-1])
                npsi[I][J] = (1. - dt * ta) * psi[I][J] + dt * c * (sg - ta) / (2 * hy) * (uu[1:-1, 2:] - uu[1:-1, :-2])
        for k in range(len(srci)):                    # MPIAcoustic.jl:376-381: local scatter_add after the step
            gi, gj = int(srci[k]) - 1, int(srcj[k]) - 1
            unew[gi // n][gj // n][gi % n, gj % n] += srcv[s - 1, k] * dt ** 2
        wold, w, phi, psi = w, unew, nphi, npsi
        for I in range(M):
            for J in range(N):
                u_hist[s, I * n:(I + 1) * n, J * n:(J + 1) * n] = unew[I][J]
    return u_hist
