"""Edge cases through the C ABI against the oracle: empty source / receiver sets, the shortest time loops, grids
smaller than one tile or made of absorbing frame only, no absorbing frame at all, srcv longer than NSTEP, ignored
elastic point types (AddSource.cpp:83, GetReceive.cpp:43)."""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _ac(po, A, ctx, NX, NY, NSTEP, srci, srcj, srcv, rcvi, rcvj, npml=4, use=(True,) * 4, kernel=1, seed=0):
    rng = np.random.default_rng(seed)
    dx, dy, dt, vp = 10.0, 9.0, 1e-3, 2500.0
    sig, tau = po.acoustic_pml(NX, NY, dx, dy, npml=npml, vp_ref=vp, use=use)
    c = vp * (1 + 0.1 * rng.random((NX + 2, NY + 2)))
    if kernel == 0:
        u0, up0, r0 = po.acoustic_forward(NX, NY, NSTEP, dt, dx, dy, sig, tau, c, srci, srcj, srcv, rcvi, rcvj, kernel=0)
    else:
        u0, r0 = po.acoustic_forward(NX, NY, NSTEP, dt, dx, dy, sig, tau, c, srci, srcj, srcv, rcvi, rcvj)
        up0 = None
    obs = 0.5 * r0 + 0.1 * rng.standard_normal(r0.shape)
    L0, g0, s0 = po.acoustic_misfit_grad(NX, NY, NSTEP, dt, dx, dy, sig, tau, c, srci, srcj, rcvi, rcvj, obs, u0,
                                         upre_hist=up0)
    p = A.AcousticPropagatorParams(NX=NX, NY=NY, NSTEP=NSTEP, DELTAX=dx, DELTAY=dy, DELTAT=dt, vp_ref=vp,
                                   NPOINTS_PML=npml, PropagatorKernel=kernel, USE_PML_XMIN=use[0], USE_PML_XMAX=use[1],
                                   USE_PML_YMIN=use[2], USE_PML_YMAX=use[3])
    R = A.acoustic_misfit_grad(p, A.AcousticSource(srci, srcj, srcv), c, A.AcousticReceiver(rcvi, rcvj), obs, ctx=ctx)
    if kernel == 1:
        assert np.array_equal(R["rcvv"], r0)
    else:
        assert relerr(R["rcvv"], r0) < 1e-12
    assert abs(R["loss"] - L0) <= 1e-12 * max(L0, 1e-300)
    assert relerr(R["grad_c"], g0) < TOL and relerr(R["grad_srcv"], s0) < TOL
    return R, u0


@pytest.mark.parametrize("kernel", [1, 0])
def test_acoustic_no_receivers_no_sources(A, ctx, po, kernel):
    srcv = np.stack([po.ricker(30, 5.0, 8.0, 1e6)], 1)
    R, _ = _ac(po, A, ctx, 40, 50, 30, [20], [25], srcv, [], [], kernel=kernel)          # no receivers
    assert R["rcvv"].shape == (31, 0) and R["loss"] == 0.0 and not R["grad_c"].any()
    R, u0 = _ac(po, A, ctx, 40, 50, 30, [], [], np.zeros((30, 0)), [3, 7], [4, 9], kernel=kernel)   # no sources
    assert not u0.any() and R["grad_srcv"].shape == (30, 0) and R["loss"] > 0 and not R["grad_c"].any()


@pytest.mark.parametrize("kernel", [1, 0])
@pytest.mark.parametrize("NSTEP", [2, 3])
def test_acoustic_shortest_loops(A, ctx, po, NSTEP, kernel):
    srcv = np.arange(1.0, 2 * NSTEP + 1).reshape(NSTEP, 2) * 1e6
    _ac(po, A, ctx, 30, 600, NSTEP, [10, 1], [300, 7], srcv, [10, 11, 2], [300, 300, 8], kernel=kernel)


@pytest.mark.parametrize("kernel", [1, 0])
@pytest.mark.parametrize("shape", [(3, 3), (5, 700), (700, 4), (9, 9)])
def test_acoustic_tiny_and_frame_only_grids(A, ctx, po, shape, kernel):
    NX, NY = shape
    srcv = np.stack([po.ricker(25, 5.0, 8.0, 1e6), po.ricker(25, 4.0, 6.0, 2e6)], 1)
    _ac(po, A, ctx, NX, NY, 25, [2, NX + 1], [2, NY], srcv, [1, 2, NX + 2, 3], [1, 3, NY + 2, 2], npml=4, kernel=kernel)


@pytest.mark.parametrize("kernel", [1, 0])
def test_acoustic_no_absorbing_frame_and_long_srcv(A, ctx, po, kernel):
    srcv = np.stack([po.ricker(60, 6.0, 9.0, 1e6)], 1)          # 60 rows, NSTEP = 40: the extra rows are ignored
    R, _ = _ac(po, A, ctx, 70, 530, 40, [35], [260], srcv, np.arange(5, 60), np.full(55, 200), use=(False,) * 4,
               kernel=kernel)
    assert R["grad_srcv"].shape[0] == 40


def test_elastic_edge_cases(A, ctx, po):
    rng = np.random.default_rng(4)
    for variant, NX, NY, NSTEP in ((0, 6, 6, 1), (1, 6, 7, 2), (0, 40, 300, 5)):
        H, W = po.elastic_dims(variant, NX, NY)
        h, dt = 1.0, 1e-4
        ax, bx = po.elastic_cpml_1d(NX, h, dt, npml=3, vp_ref=3300.0, alpha_max=np.pi * 15)
        ay, by = po.elastic_cpml_1d(NY, h, dt, npml=3, vp_ref=3300.0, alpha_max=np.pi * 15)
        vp = 3000.0 * (1 + 0.1 * rng.random((H, W)))
        vs, rho = vp / 1.732, 2800.0 * (1 + 0.1 * rng.random((H, W)))
        mu, lam = rho * vs * vs, rho * (vp * vp - 2 * vs * vs)
        srci, srcj = np.array([3, 4, 3]), np.array([3, 3, 4])
        srctype = np.array([2, 7, 0])                     # type 7 is ignored (AddSource.cpp:83)
        srcv = rng.standard_normal((NSTEP + 3, 3)) * 1e3  # more rows than NSTEP
        for rcvi, rcvj, rcvtype in (([], [], []), ([2, 3, 4], [3, 3, 2], [0, 9, 4])):   # none / one ignored type
            args = (variant, NX, NY, NSTEP, dt, h, h, ax, bx, ay, by, rho, lam, mu, srci, srcj, srctype, srcv, rcvi, rcvj,
                    rcvtype)
            r0, _ = po.elastic_forward(*args)
            obs = 0.5 * r0
            O = po.elastic_misfit_grad(*args, obs)
            p = A.ElasticPropagatorParams(NX=NX, NY=NY, NSTEP=NSTEP, DELTAX=h, DELTAY=h, DELTAT=dt, NPOINTS_PML=3,
                                          vp_ref=3300.0, ALPHA_MAX_PML=np.pi * 15, variant=variant)
            unpad = (lambda a: a) if variant == 0 else (lambda a: a[2:-2, 2:-2])
            R = A.elastic_misfit_grad(p, A.ElasticSource(srci, srcj, srctype, srcv), unpad(rho), unpad(lam), unpad(mu),
                                      A.ElasticReceiver(rcvi, rcvj, rcvtype), obs, ctx=ctx)
            assert R["rcvv"].shape == (len(rcvi), NSTEP + 1) and np.array_equal(R["rcvv"], r0)
            assert abs(R["loss"] - O["loss"]) <= 1e-12 * max(O["loss"], 1e-300)
            if len(rcvi):
                assert relerr(R["grad_srcv"], O["grad_srcv"]) < TOL
                for k, gk in (("grad_rho", "grad_rho"), ("grad_lambda", "grad_lam"), ("grad_mu", "grad_mu")):
                    assert relerr(R[k], unpad(O[gk])) < TOL, (variant, NX, k)
            else:
                assert R["loss"] == 0.0 and not R["grad_mu"].any() and not R["grad_srcv"].any()
