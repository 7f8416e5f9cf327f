"""Multi-GPU worker (launched by torchrun from tests/test_multigpu.py or by hand):
   torchrun --nproc-per-node N --master-addr 127.0.0.1 tests/mgpu_worker.py
Checks on every rank: slab-decomposed acoustic forward+gradient == undecomposed CPU oracle (traces bit-identical,
gradients <= 1e-10; the reference's own distributed test is decomposed == undecomposed,
examples/mpi_acoustic/verification/verify_forward.jl:78-95), with and without checkpoint segments, and the
shot-parallel loss/gradient == the sum over shots."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import adseis_b200 as A  # noqa: E402
from adseis_b200 import parallel  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def relerr(a, b):
    d = np.abs(b).max()
    return float(np.abs(a - b).max() / d) if d > 0 else float(np.abs(a - b).max())


def main():
    rank, world, local_rank = parallel.init_process_group("nccl")
    ctx = A.Context(local_rank)
    rng = np.random.default_rng(42)                       # same inputs on every rank
    # ---------------- domain decomposition ----------------
    NX, NY, NSTEP, dx, dt, vp = 97, 600, 60, 10.0, 1e-3, 2500.0
    sig, tau = po.acoustic_pml(NX, NY, dx, dx, npml=8, vp_ref=vp)
    c = vp * (1 + 0.1 * rng.random((NX + 2, NY + 2)))
    nsrc, nrcv = 5, 60
    srci = rng.integers(1, NX + 2, nsrc); srcj = rng.integers(2, NY, nsrc)
    rcvi = rng.integers(1, NX + 2, nrcv); rcvj = rng.integers(2, NY, nrcv)
    # put sources and receivers on both sides of every slab boundary
    bounds = [parallel.slab_partition(NX, world, r) for r in range(world)]
    for k, (r0, r1) in enumerate(bounds[:-1]):
        srci[k % nsrc] = r1            # padded row r1-1 (last row of slab k), 1-based
        rcvi[2 * k] = r1               # last row of slab k
        rcvi[2 * k + 1] = r1 + 1       # first row of slab k+1
    srcv = np.stack([po.ricker(NSTEP, 8.0 + k, 20.0 + 2 * k, 1e6) for k in range(nsrc)], 1)
    u0, r0 = po.acoustic_forward(NX, NY, NSTEP, dt, dx, dx, sig, tau, c, srci, srcj, srcv, rcvi, rcvj)
    obs = 0.7 * r0 + 0.02 * np.abs(r0).max() * rng.standard_normal(r0.shape)
    L0, g0, s0 = po.acoustic_misfit_grad(NX, NY, NSTEP, dt, dx, dx, sig, tau, c, srci, srcj, rcvi, rcvj, obs, u0)
    p = A.AcousticPropagatorParams(NX=NX, NY=NY, NSTEP=NSTEP, DELTAX=dx, DELTAY=dx, DELTAT=dt, vp_ref=vp,
                                   NPOINTS_PML=8)
    for slots in (None, 14):          # full history, then a 14-snapshot window (checkpointed reverse sweep)
        dd = parallel.DomainDecomposedAcoustic(p, srci, srcj, rcvi, rcvj, ctx=ctx, hist_slots=slots)
        dd.set_model(c); dd.set_srcv(srcv); dd.set_obs(obs)
        dd.forward()
        r = dd.rcvv()
        assert np.array_equal(r, r0), "rank %d: DD traces differ (max %g)" % (rank, np.abs(r - r0).max())
        dd.gradient()
        L, g, s = dd.loss(), dd.grad_c().cpu().numpy(), dd.grad_srcv()
        assert abs(L - L0) / L0 < 1e-12, (L, L0)
        assert relerr(g, g0) < 1e-10, relerr(g, g0)
        assert relerr(s, s0) < 1e-10, relerr(s, s0)
        info = dd.plan.info()
        if slots:
            assert info["segments"] > 1 and info["recomputed_steps"] > 0
        if rank == 0:
            print("DD x%d slots=%s ok: loss rel %.1e grad_c rel %.1e grad_srcv rel %.1e segments %d" %
                  (world, slots, abs(L - L0) / L0, relerr(g, g0), relerr(s, s0), info["segments"]), flush=True)
        dd.close()
    # ---------------- shot parallelism ----------------
    nshots = 5
    NX, NY, NSTEP = 60, 300, 40
    sig, tau = po.acoustic_pml(NX, NY, dx, dx, npml=8, vp_ref=vp)
    c = vp * (1 + 0.1 * rng.random((NX + 2, NY + 2)))
    p = A.AcousticPropagatorParams(NX=NX, NY=NY, NSTEP=NSTEP, DELTAX=dx, DELTAY=dx, DELTAT=dt, vp_ref=vp,
                                   NPOINTS_PML=8)
    srcs, rcvs, Rs, Ltot, gtot = [], [], [], 0.0, 0.0
    for k in range(nshots):
        si, sj = np.array([10 + 8 * k]), np.array([40 + 50 * k])
        sv = po.ricker(NSTEP, 7.0, 15.0, 1e6).reshape(-1, 1)
        ri, rj = np.full(30, 5), np.arange(20, 290, 9)
        u, r = po.acoustic_forward(NX, NY, NSTEP, dt, dx, dx, sig, tau, c, si, sj, sv, ri, rj)
        ob = 0.5 * r
        L, g, _ = po.acoustic_misfit_grad(NX, NY, NSTEP, dt, dx, dx, sig, tau, c, si, sj, ri, rj, ob, u)
        Ltot += L; gtot = gtot + g
        srcs.append(A.AcousticSource(si, sj, sv)); rcvs.append(A.AcousticReceiver(ri, rj)); Rs.append(ob)
    L, g = parallel.compute_loss_and_grads_GPU(p, srcs, rcvs, Rs, c, ctx=ctx)
    assert abs(L - Ltot) / Ltot < 1e-12 and relerr(g, gtot) < 1e-10
    if rank == 0:
        print("shot-parallel x%d ok: loss rel %.1e grad rel %.1e" % (world, abs(L - Ltot) / Ltot, relerr(g, gtot)), flush=True)
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
