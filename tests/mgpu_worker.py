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
    p = A.AcousticPropagatorParams(PropagatorKernel=1, NX=NX, NY=NY, NSTEP=NSTEP, DELTAX=dx, DELTAY=dx, DELTAT=dt, vp_ref=vp,
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
    # ---------------- the halo protocol is an execution detail: packed rows == fence + flag, run to run, bit for bit ----------------
    # (a taller grid so that the two-step path with its frame-only launches is exercised on slabs as well)
    NXb = 64 * world + 30
    sigb, taub = po.acoustic_pml(NXb, NY, dx, dx, npml=8, vp_ref=vp)
    cb = vp * (1 + 0.1 * rng.random((NXb + 2, NY + 2)))
    pb = A.AcousticPropagatorParams(PropagatorKernel=1, NX=NXb, NY=NY, NSTEP=NSTEP, DELTAX=dx, DELTAY=dx, DELTAT=dt, vp_ref=vp,
                                    NPOINTS_PML=8)
    bnd = [parallel.slab_partition(NXb, world, r) for r in range(world)]
    sib = rng.integers(1, NXb + 2, nsrc); rib = rng.integers(1, NXb + 2, nrcv)
    for k, (r0_, r1_) in enumerate(bnd[:-1]):
        sib[k % nsrc] = r1_ + (k & 1); rib[2 * k] = r1_; rib[2 * k + 1] = r1_ + 1
    ub, rb_ = po.acoustic_forward(NXb, NY, NSTEP, dt, dx, dx, sigb, taub, cb, sib, srcj, srcv, rib, rcvj)
    obsb = 0.6 * rb_
    Lb, gb, sb_ = po.acoustic_misfit_grad(NXb, NY, NSTEP, dt, dx, dx, sigb, taub, cb, sib, srcj, rib, rcvj, obsb, ub)
    got = {}
    for tb in ("0", "1"):
        for ll in ("1", "0"):
            os.environ["ADSEIS_AC_TB_SLAB"] = tb; os.environ["ADSEIS_AC_TB_SLAB_MIN"] = "0"; os.environ["ADSEIS_AC_LL"] = ll
            for rep in range(2):
                dd = parallel.DomainDecomposedAcoustic(pb, sib, srcj, rib, rcvj, ctx=ctx, hist_slots=None if rep == 0 else 14)
                dd.set_model(cb); dd.set_srcv(srcv); dd.set_obs(obsb)
                dd.gradient()
                got[(tb, ll, rep)] = (dd.rcvv(), dd.loss(), dd.grad_c().cpu().numpy(), dd.grad_srcv())
                dd.close()
    for k in ("ADSEIS_AC_TB_SLAB", "ADSEIS_AC_TB_SLAB_MIN", "ADSEIS_AC_LL"):
        os.environ.pop(k)
    ref_ = got[("0", "0", 0)]
    assert np.array_equal(ref_[0], rb_) and relerr(ref_[2], gb) < 1e-10 and relerr(ref_[3], sb_) < 1e-10
    for key, val in got.items():
        assert np.array_equal(val[0], ref_[0]) and val[1] == ref_[1] and np.array_equal(val[2], ref_[2]) and \
            np.array_equal(val[3], ref_[3]), "slab run %s differs from the fence+flag one-step run" % (key,)
    if rank == 0:
        print("DD x%d halo protocols ok: one-step / two-step x packed / flags x resident / checkpointed: identical bits" % world,
              flush=True)
    # ---------------- PropagatorKernel = 0 on slabs (MPIAcousticPropagatorSolver's scheme, MPIAcoustic.jl:212-246) ----------------
    # phi', psi' are driven by the new wavefield: two halo rows, explicit exchange after every step launch.  Sources sit
    # on both sides of every slab boundary (their injected part is removed from the NEIGHBOUR's c-gradient terms too).
    srci0, srcj0 = srci.copy(), srcj.copy()
    for k, (r0_, r1_) in enumerate(bounds[:-1]):
        srci0[(2 * k) % nsrc] = r1_; srci0[(2 * k + 1) % nsrc] = r1_ + 1
        srcj0[(2 * k) % nsrc] = 3 + k; srcj0[(2 * k + 1) % nsrc] = NY - 2 - k      # inside the absorbing frame columns
    u0k, up0k, r0k = po.acoustic_forward(NX, NY, NSTEP, dt, dx, dx, sig, tau, c, srci0, srcj0, srcv, rcvi, rcvj, kernel=0)
    obsk = 0.7 * r0k + 0.02 * np.abs(r0k).max() * rng.standard_normal(r0k.shape)
    L0k, g0k, s0k = po.acoustic_misfit_grad(NX, NY, NSTEP, dt, dx, dx, sig, tau, c, srci0, srcj0, rcvi, rcvj, obsk, u0k,
                                            upre_hist=up0k)
    p0 = A.AcousticPropagatorParams(PropagatorKernel=0, NX=NX, NY=NY, NSTEP=NSTEP, DELTAX=dx, DELTAY=dx, DELTAT=dt, vp_ref=vp,
                                    NPOINTS_PML=8)
    one = A.AcousticPlan(p0, srci0, srcj0, rcvi, rcvj, ctx=ctx)            # the undecomposed CUDA path, on every rank
    one.set_model(c); one.set_srcv(srcv); one.set_obs(obsk); one.gradient()
    r1k, g1k, s1k = one.rcvv(), one.grad_c(), one.grad_srcv()
    one.close()
    assert relerr(r1k, r0k) < 1e-12
    for slots in (None, 14):
        dd = parallel.DomainDecomposedAcoustic(p0, srci0, srcj0, rcvi, rcvj, ctx=ctx, hist_slots=slots)
        dd.set_model(c); dd.set_srcv(srcv); dd.set_obs(obsk)
        dd.forward()
        r = dd.rcvv()
        assert np.array_equal(r, r1k), "rank %d: PropagatorKernel=0 DD traces differ from the undecomposed run (max %g)" % (
            rank, np.abs(r - r1k).max())
        dd.gradient()
        L, g, s_ = dd.loss(), dd.grad_c().cpu().numpy(), dd.grad_srcv()
        assert abs(L - L0k) / L0k < 1e-12, (L, L0k)
        assert relerr(g, g0k) < 1e-10, relerr(g, g0k)
        assert relerr(s_, s0k) < 1e-10, relerr(s_, s0k)
        assert relerr(g, g1k) < 1e-13 and relerr(s_, s1k) < 1e-13, (relerr(g, g1k), relerr(s_, s1k))
        if rank == 0:
            print("DD x%d PropagatorKernel=0 slots=%s ok: grad_c rel %.1e grad_srcv rel %.1e segments %d" %
                  (world, slots, relerr(g, g0k), relerr(s_, s0k), dd.plan.info()["segments"]), flush=True)
        dd.close()
    # ---------------- elastic domain decomposition (both reference variants) ----------------
    # the reference's own test is decomposed == undecomposed (examples/mpi_elastic/verification/verify_backward.jl)
    for variant, (NX, NY, NSTEP) in ((1, (36 * world + 3, 150, 24)), (0, (34 * world + 1, 140, 22))):
        rng2 = np.random.default_rng(7 + variant)
        h, dte, npml = 1.0, 1e-4, 8
        H, W = po.elastic_dims(variant, NX, NY)
        ax, bx = po.elastic_cpml_1d(NX, h, dte, npml=npml, vp_ref=3300.0, alpha_max=np.pi * 15)
        ay, by = po.elastic_cpml_1d(NY, h, dte, npml=npml, vp_ref=3300.0, alpha_max=np.pi * 15)
        vpm = 3000.0 * (1 + 0.1 * rng2.random((H, W)))
        vsm = vpm / 1.732 * (1 + 0.05 * rng2.random((H, W)))
        rho = 2800.0 * (1 + 0.1 * rng2.random((H, W)))
        mu, lam = rho * vsm * vsm, rho * (vpm * vpm - 2 * vsm * vsm)
        pe = A.ElasticPropagatorParams(NX=NX, NY=NY, NSTEP=NSTEP, DELTAX=h, DELTAY=h, DELTAT=dte, NPOINTS_PML=npml,
                                       vp_ref=3300.0, ALPHA_MAX_PML=np.pi * 15, variant=variant)
        nsrc, nrcv = 8, 40
        srci = rng2.integers(3, NX - 2, nsrc); srcj = rng2.integers(3, NY - 2, nsrc)
        srctype = np.arange(nsrc) % 5
        rcvi = rng2.integers(1, NX + 1, nrcv); rcvj = rng2.integers(1, NY + 1, nrcv)
        rcvtype = rng2.integers(0, 5, nrcv)
        # sources and receivers of every type on the rows next to every slab boundary
        ioff = 1 if variant == 0 else -1     # internal 0-based row -> the caller's 1-based index
        for r in range(world - 1):
            b = parallel.elastic_slab_partition(pe, world, r)[1]    # first internal row of slab r+1
            for k in range(5):
                rcvi[(10 * r + 2 * k) % nrcv] = b - 1 + ioff; rcvtype[(10 * r + 2 * k) % nrcv] = k
                rcvi[(10 * r + 2 * k + 1) % nrcv] = b + ioff; rcvtype[(10 * r + 2 * k + 1) % nrcv] = k
            srci[(2 * r) % nsrc] = b - 1 + ioff
            srci[(2 * r + 1) % nsrc] = b + ioff
        srcv = np.stack([po.ricker(NSTEP, 6.0 + k, 10.0 + k, 1e3 * (1 + k)) for k in range(nsrc)], 1)
        args = (variant, NX, NY, NSTEP, dte, h, h, ax, bx, ay, by, rho, lam, mu, srci, srcj, srctype, srcv, rcvi, rcvj,
                rcvtype)
        r0, _ = po.elastic_forward(*args)
        obs = 0.6 * r0 + 0.05 * np.abs(r0).max() * rng2.standard_normal(r0.shape)
        O = po.elastic_misfit_grad(*args, obs)
        unpad = (lambda a: a) if variant == 0 else (lambda a: a[2:-2, 2:-2])
        for slots in (None, 7):
            # full history: TMA-ring marching CTAs forced onto the small box; checkpointed: generic CTAs only
            os.environ["ADSEIS_EL_MARCH_MIN"] = "0" if slots is None else str(1 << 40)
            dd = parallel.DomainDecomposedElastic(pe, srci, srcj, srctype, rcvi, rcvj, rcvtype, ctx=ctx, hist_slots=slots)
            dd.set_model(unpad(rho), unpad(lam), unpad(mu)); dd.set_srcv(srcv); dd.set_obs(obs)
            dd.forward()
            r = dd.rcvv()
            assert np.array_equal(r, r0), "rank %d: elastic DD traces differ (max %g)" % (rank, np.abs(r - r0).max())
            dd.gradient(True)
            L, gs = dd.loss(), dd.grad_srcv()
            gr, gl, gm = dd.grads()
            assert abs(L - O["loss"]) / O["loss"] < 1e-12, (L, O["loss"])
            errs = dict(srcv=relerr(gs, O["grad_srcv"]), rho=relerr(gr, unpad(O["grad_rho"])),
                        lam=relerr(gl, unpad(O["grad_lam"])), mu=relerr(gm, unpad(O["grad_mu"])))
            assert max(errs.values()) < 1e-10, errs
            info = dd.plan.info()
            dd.gradient(False)                      # source-time-function gradient: no forward history at all
            assert relerr(dd.grad_srcv(), O["grad_srcv"]) < 1e-10
            if slots:
                assert info["segments"] > 1 and info["recomputed_steps"] > 0
            if rank == 0:
                print("elastic DD x%d variant %d slots=%s ok: %s segments %d" % (world, variant, slots,
                      " ".join("%s %.1e" % kv for kv in errs.items()), info["segments"]), flush=True)
            dd.close()
    # ---------------- shot parallelism ----------------
    nshots = 5
    NX, NY, NSTEP = 60, 300, 40
    sig, tau = po.acoustic_pml(NX, NY, dx, dx, npml=8, vp_ref=vp)
    c = vp * (1 + 0.1 * rng.random((NX + 2, NY + 2)))
    p = A.AcousticPropagatorParams(PropagatorKernel=1, NX=NX, NY=NY, NSTEP=NSTEP, DELTAX=dx, DELTAY=dx, DELTAT=dt, vp_ref=vp,
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
    # one plan per GPU re-pointed at every shot (adseis_acoustic_plan_set_points), reused across evaluations
    cache = parallel.ShotPlanCache()
    R2 = parallel.compute_forward_GPU(p, srcs, rcvs, c, ctx=ctx, plan_cache=cache)       # compute_forward_GPU, Utils.jl:574-600
    assert all(np.array_equal(R2[k], 2.0 * Rs[k]) for k in range(nshots))               # obs were 0.5 * traces
    for _ in range(2):
        L2, g2 = parallel.compute_loss_and_grads_GPU(p, srcs, rcvs, Rs, c, ctx=ctx, plan_cache=cache)
        assert L2 == L and np.array_equal(g2, g)
    cache.close()
    # elastic sources through the same entry point (the reference's signature takes ElasticSource too, Utils.jl:300)
    NXe, NYe, NSe = 50, 64, 24
    axe, bxe = po.elastic_cpml_1d(NXe, 1.0, 1e-4, npml=6, vp_ref=3300.0, alpha_max=np.pi * 15)
    aye, bye = po.elastic_cpml_1d(NYe, 1.0, 1e-4, npml=6, vp_ref=3300.0, alpha_max=np.pi * 15)
    rng3 = np.random.default_rng(5)
    rho = 2800.0 * (1 + 0.1 * rng3.random((NXe + 2, NYe + 2)))
    vpe = 3000.0 * (1 + 0.1 * rng3.random((NXe + 2, NYe + 2)))
    mu, lam = rho * (vpe / 1.732) ** 2, rho * (vpe ** 2 - 2 * (vpe / 1.732) ** 2)
    pe = A.ElasticPropagatorParams(NX=NXe, NY=NYe, NSTEP=NSe, DELTAX=1.0, DELTAY=1.0, DELTAT=1e-4, NPOINTS_PML=6,
                                   vp_ref=3300.0, ALPHA_MAX_PML=np.pi * 15, variant=0)
    es, er, eR, eL, eg = [], [], [], 0.0, [0.0, 0.0, 0.0]
    for k in range(3):
        si, sj, ty = np.array([12 + 9 * k]), np.array([20 + 10 * k]), np.array([k])
        sv = po.ricker(NSe, 5.0, 8.0, 1e4).reshape(-1, 1)
        ri, rj, rt = np.arange(8, 44, 3), np.full(12, 30), np.arange(12) % 5
        a = (0, NXe, NYe, NSe, 1e-4, 1.0, 1.0, axe, bxe, aye, bye, rho, lam, mu, si, sj, ty, sv, ri, rj, rt)
        r0e, _ = po.elastic_forward(*a)
        O = po.elastic_misfit_grad(*a, 0.5 * r0e)
        eL += O["loss"]
        for q, key in enumerate(("grad_rho", "grad_lam", "grad_mu")):
            eg[q] = eg[q] + O[key]
        es.append(A.ElasticSource(si, sj, ty, sv)); er.append(A.ElasticReceiver(ri, rj, rt)); eR.append(0.5 * r0e)
    Le, ge = parallel.compute_loss_and_grads_GPU(pe, es, er, eR, (rho, lam, mu), ctx=ctx)
    assert abs(Le - eL) / eL < 1e-12 and all(relerr(ge[q], eg[q]) < 1e-10 for q in range(3))
    if rank == 0:
        print("shot-parallel x%d ok: loss rel %.1e grad rel %.1e; plan reuse and elastic shots ok" %
              (world, abs(L - Ltot) / Ltot, relerr(g, gtot)), flush=True)
    import torch.distributed as dist
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
