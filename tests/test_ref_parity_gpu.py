"""GPU parity, closing the chain in ONE run on the GPU box:
  (1) CUDA path (through the C ABI) vs the graph over the REFERENCE'S OWN op bodies (oracle/_ref, ref_graph.inc):
      elastic S / M and acoustic PropagatorKernel=0, single process and block-decomposed MPI emulation;
  (2) the op-level C-ABI entry points adseis_op_add_source_fwd / adseis_op_get_receive_fwd vs the reference's
      AddSource.cpp / GetReceive.cpp bodies;
  (3) BASELINE.json's C1 (acoustic 401x133) and C2 (elastic 500^2) at their full step counts (nt=1000) vs the CPU oracle.
Bars: forward traces / fields bit-identical where the reference's path is the same IEEE expression (elastic, acoustic
custom op), 1e-12 for PropagatorKernel=0 forward (TF element-wise kernels vs our fused expression), gradients <= 1e-10
relative (BASELINE.json north_star)."""
import ctypes as C

import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu
TOL = 1e-10


@pytest.fixture(autouse=True)
def _need_ref(po):
    if not po.has_ref():
        pytest.skip("oracle/_ref not built")


@pytest.fixture(params=["march", "generic"])
def tiling(request, monkeypatch):
    monkeypatch.setenv("ADSEIS_EL_MARCH_MIN", "0" if request.param == "march" else str(1 << 40))
    return request.param


def _unpad(variant, a):
    return a if variant == 0 else a[..., 2:-2, 2:-2]


@pytest.mark.parametrize("variant,NX,NY,block", [(0, 96, 300, None), (1, 90, 280, None), (1, 96, 288, (48, 96))])
def test_elastic_cuda_vs_reference_op_graph(A, ctx, po, tiling, variant, NX, NY, block):
    rng = np.random.default_rng(7 * NX + variant)
    NSTEP, h, dt, npml = 18, 1.0, 1e-4, 8
    H, W = po.elastic_dims(variant, NX, NY)
    ax, bx = po.elastic_cpml_1d(NX, h, dt, npml=npml, vp_ref=3300.0, alpha_max=np.pi * 15)
    ay, by = po.elastic_cpml_1d(NY, h, dt, npml=npml, vp_ref=3300.0, alpha_max=np.pi * 15)
    vp = 3000.0 * (1 + 0.1 * rng.random((H, W)))
    vs = vp / 1.732 * (1 + 0.05 * rng.random((H, W)))
    rho = 2800.0 * (1 + 0.1 * rng.random((H, W)))
    mu, lam = rho * vs * vs, rho * (vp * vp - 2 * vs * vs)
    nsrc = 8
    srci, srcj = rng.integers(3, NX - 2, nsrc), rng.integers(3, NY - 2, nsrc)
    srctype = np.array([0, 1, 2, 3, 4, 2, 0, 1])
    srci[5], srcj[5] = srci[2], srcj[2]
    srci[6], srcj[6] = 2, 2                       # inside the CPML corner
    srci[7], srcj[7] = NX // 2, NY // 2           # marching box
    srcv = np.stack([po.ricker(NSTEP, 4.0 + k, 6.0 + k, 1e3 * (1 + k)) for k in range(nsrc)], 1)
    nrcv = 40
    rcvi, rcvj, rcvtype = rng.integers(1, NX + 1, nrcv), rng.integers(1, NY + 1, nrcv), rng.integers(0, 5, nrcv)
    rcvi[0], rcvj[0], rcvtype[0] = srci[2], srcj[2], 2
    rcvi[1], rcvj[1], rcvtype[1] = rcvi[0], rcvj[0], 2
    args = (variant, NX, NY, NSTEP, dt, h, h, ax, bx, ay, by, rho, lam, mu, srci, srcj, srctype, srcv, rcvi, rcvj, rcvtype)
    r0 = po.ref_elastic(*args, want_grad=False, block=block)["rcvv"]
    obs = 0.6 * r0 + 0.05 * np.abs(r0).max() * rng.standard_normal(r0.shape)
    R = po.ref_elastic(*args, obs, want_hist=True, block=block)
    p = A.ElasticPropagatorParams(NX=NX, NY=NY, NSTEP=NSTEP, DELTAX=h, DELTAY=h, DELTAT=dt, NPOINTS_PML=npml,
                                  vp_ref=3300.0, ALPHA_MAX_PML=np.pi * 15, variant=variant)
    plan = A.ElasticPlan(p, srci, srcj, srctype, rcvi, rcvj, rcvtype, ctx=ctx)
    plan.set_model(*(np.ascontiguousarray(_unpad(variant, x)) for x in (rho, lam, mu)))
    plan.set_srcv(srcv); plan.set_obs(obs)
    plan.gradient(True)
    assert np.abs(r0).max() > 0 and np.array_equal(plan.rcvv(), R["rcvv"])          # bit for bit
    for f in range(5):
        for s in (1, NSTEP // 2, NSTEP):
            assert np.array_equal(plan.snapshot(f, s), _unpad(variant, R["hist"][f, s])), (f, s)
    assert abs(plan.loss() - R["loss"]) <= 1e-12 * R["loss"]
    assert relerr(plan.grad_srcv(), R["grad_srcv"]) < TOL
    assert relerr(plan.grad_rho(), _unpad(variant, R["grad_rho"])) < TOL
    assert relerr(plan.grad_lambda(), _unpad(variant, R["grad_lam"])) < TOL
    assert relerr(plan.grad_mu(), _unpad(variant, R["grad_mu"])) < TOL
    plan.close()


@pytest.mark.parametrize("mpi,shape,block", [(False, (70, 560, 40), None), (True, (72, 576, 36), (36, 192)),
                                             (True, (60, 40, 50), (20, 20))])
def test_acoustic_kernel0_cuda_vs_reference_op_graph(A, ctx, po, mpi, shape, block):
    """PropagatorKernel=0 (Core.jl:528-549 / MPIAcoustic.jl:212-246 incl. the exchange of the new wavefield)."""
    NX, NY, NSTEP = shape
    rng = np.random.default_rng(NX + NY)
    dx, dy, dt, vp, npml = 10.0, 8.0, 1e-3, 2500.0, 8
    sig, tau = po.acoustic_pml(NX, NY, dx, dy, npml=npml, vp_ref=vp)
    nsrc, nrcv = 5, 30
    off = 0 if mpi else 1           # padded 1-based indices address [1, N+2]; the MPI convention [1, N]
    srci = np.array([NX // 2, 3, 4, 1, NX - 1]) + off
    srcj = np.array([NY // 2, 5, 5, NY // 3, NY]) + off
    srcv = np.stack([po.ricker(NSTEP, 6.0 + k, 10.0 + k, 1e6) for k in range(nsrc)], 1)
    rcvi, rcvj = rng.integers(1, NX + 1, nrcv) + off, rng.integers(1, NY + 1, nrcv) + off
    if mpi:
        c2 = (vp * (1 + 0.1 * rng.random((NX, NY)))) ** 2
        run = lambda obs, **kw: po.ref_mpi_acoustic_graph(0, NX, NY, block, NSTEP, dt, dx, dy, sig, tau, c2, srci, srcj,
                                                          srcv, rcvi, rcvj, obs, **kw)
        model, gkey = c2, "grad_c2"
    else:
        c = vp * (1 + 0.1 * rng.random((NX + 2, NY + 2)))
        run = lambda obs, **kw: po.ref_acoustic_graph(0, NX, NY, NSTEP, dt, dx, dy, sig, tau, c, srci, srcj, srcv, rcvi,
                                                      rcvj, obs, **kw)
        model, gkey = c, "grad_c"
    r0 = run(None, want_grad=False)["rcvv"]
    obs = 0.7 * r0 + 0.02 * np.abs(r0).max() * rng.standard_normal(r0.shape)
    R = run(obs)
    p = A.AcousticPropagatorParams(NX=NX, NY=NY, NSTEP=NSTEP, DELTAX=dx, DELTAY=dy, DELTAT=dt, vp_ref=vp,
                                   NPOINTS_PML=npml, PropagatorKernel=0, mpi_convention=mpi)
    G = A.acoustic_misfit_grad(p, A.AcousticSource(srci, srcj, srcv), model, A.AcousticReceiver(rcvi, rcvj), obs, ctx=ctx)
    assert np.abs(r0).max() > 0 and relerr(G["rcvv"], R["rcvv"]) < 1e-12
    assert abs(G["loss"] - R["loss"]) <= 1e-12 * R["loss"]
    assert relerr(G["grad_c"], R[gkey]) < TOL and relerr(G["grad_srcv"], R["grad_srcv"]) < TOL


def test_op_level_add_source_and_get_receive_vs_reference_bodies(A, ctx, po):
    """adseis_op_add_source_fwd / adseis_op_get_receive_fwd on device pointers == AddSource.cpp:33-87 /
    GetReceive.cpp:10-46 (the reference's bodies, compiled in place into oracle/_ref), gradtest.jl-style inputs."""
    import torch
    lib, ref = A._lib.load(), po.ref_lib()
    rng = np.random.default_rng(233)
    NX, NY, nt = 23, 31, 7
    N = (NX + 2) * (NY + 2)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_longlong)
    fields = [rng.random(N) for _ in range(5)]
    nsrc = 11
    srci, srcj = rng.integers(1, NX + 3, nsrc), rng.integers(1, NY + 3, nsrc)
    srctype = rng.integers(0, 5, nsrc)
    srci[3], srcj[3], srctype[3] = srci[1], srcj[1], srctype[1]          # duplicates accumulate in order
    srcv = rng.standard_normal(nsrc)
    want = [np.empty(N) for _ in range(5)]
    ref.ref_op_add_source_fwd(*[x.ctypes.data_as(dp) for x in want], *[x.ctypes.data_as(dp) for x in fields],
                              srci.ctypes.data_as(ip), srcj.ctypes.data_as(ip), srcv.ctypes.data_as(dp),
                              srctype.ctypes.data_as(ip), C.c_longlong(nsrc), C.c_longlong(NX), C.c_longlong(NY))
    dev = torch.device("cuda")
    t = lambda a: torch.tensor(a, device=dev)
    d_in, d_out = [t(x) for x in fields], [torch.empty(N, dtype=torch.float64, device=dev) for _ in range(5)]
    d_si, d_sj, d_st, d_sv = t(srci), t(srcj), t(srctype), t(srcv)
    A._lib.check(lib.adseis_op_add_source_fwd(ctx.handle, *[A._lib.ptr(x) for x in d_out], *[A._lib.ptr(x) for x in d_in],
                                              A._lib.ptr(d_si), A._lib.ptr(d_sj), A._lib.ptr(d_sv), A._lib.ptr(d_st),
                                              nsrc, NX, NY, None))
    ctx.sync()
    for a, b in zip(d_out, want):
        assert np.array_equal(a.cpu().numpy(), b)
    # get_receive over nt stacked snapshots
    hist = [rng.random(nt * N) for _ in range(5)]
    nrcv = 9
    rcvi, rcvj, rcvtype = rng.integers(1, NX + 3, nrcv), rng.integers(1, NY + 3, nrcv), rng.integers(0, 5, nrcv)
    want_r = np.zeros(nrcv * nt)
    ref.ref_op_get_receive_fwd(want_r.ctypes.data_as(dp), *[x.ctypes.data_as(dp) for x in hist], C.c_longlong(nt),
                               rcvi.ctypes.data_as(ip), rcvj.ctypes.data_as(ip), rcvtype.ctypes.data_as(ip),
                               C.c_longlong(nrcv), C.c_longlong(NX), C.c_longlong(NY))
    d_h = [t(x) for x in hist]
    d_r = torch.zeros(nrcv * nt, dtype=torch.float64, device=dev)
    d_ri, d_rj, d_rt = t(rcvi), t(rcvj), t(rcvtype)
    A._lib.check(lib.adseis_op_get_receive_fwd(ctx.handle, A._lib.ptr(d_r), *[A._lib.ptr(x) for x in d_h], nt,
                                               A._lib.ptr(d_ri), A._lib.ptr(d_rj), A._lib.ptr(d_rt), nrcv, NX, NY, None))
    ctx.sync()
    assert np.array_equal(d_r.cpu().numpy(), want_r)


# ---------------------------------------------------------------------------------------------------------------
# BASELINE.json configs[0] and [1] at their full step counts
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kernel", [1, 0])
def test_c1_full_step_count_vs_oracle(A, ctx, po, kernel):
    """C1: 401 x 133, NSTEP=1000, both acoustic schemes, traces + misfit + gradients against the CPU oracle (which is
    the reference's C++ op bodies for scheme 1 and the reference-op graph's twin for scheme 0)."""
    w = A.workloads.c1(nstep=1000, kernel=kernel)
    p, sh = w["param"], w["shots"][0]
    NX, NY, NSTEP = p.NX, p.NY, p.NSTEP
    sig, tau = po.acoustic_pml(NX, NY, p.DELTAX, p.DELTAY, npml=p.NPOINTS_PML, Rcoef=p.Rcoef, vp_ref=p.vp_ref)
    a = (NX, NY, NSTEP, p.DELTAT, p.DELTAX, p.DELTAY, sig, tau)
    pts = (sh["srci"], sh["srcj"], sh["srcv"], sh["rcvi"], sh["rcvj"])
    if kernel == 0:
        _, _, obs = po.acoustic_forward(*a, w["model_obs"], *pts, kernel=0)
        u, up, r0 = po.acoustic_forward(*a, w["model"], *pts, kernel=0)
    else:
        _, obs = po.acoustic_forward(*a, w["model_obs"], *pts)
        (u, r0), up = po.acoustic_forward(*a, w["model"], *pts), None
    L0, gc0, gs0 = po.acoustic_misfit_grad(*a, w["model"], sh["srci"], sh["srcj"], sh["rcvi"], sh["rcvj"], obs, u,
                                           upre_hist=up)
    G = A.acoustic_misfit_grad(p, A.AcousticSource(sh["srci"], sh["srcj"], sh["srcv"]), w["model"],
                               A.AcousticReceiver(sh["rcvi"], sh["rcvj"]), obs, ctx=ctx)
    assert np.abs(r0).max() > 0
    if kernel == 1:
        assert np.array_equal(G["rcvv"], r0)
    else:
        assert relerr(G["rcvv"], r0) < 1e-11
    assert abs(G["loss"] - L0) <= 1e-11 * L0
    assert relerr(G["grad_c"], gc0) < TOL and relerr(G["grad_srcv"], gs0) < TOL


def test_c2_full_step_count_vs_oracle(A, ctx, po):
    """C2: elastic 500 x 500 at NSTEP=1000: traces of the full run bit-identical to the oracle; gradients (all of rho,
    lambda, mu and the source time function) on the same grid at NSTEP=100 -- the oracle's dense 13-array tape for
    1000 steps would need 26 GB of host memory."""
    w = A.workloads.c2(nstep=1000)
    p, sh = w["param"], w["shots"][0]
    NX, NY = p.NX, p.NY
    ab = A.compute_PML_Params(p)
    ax, bx, ay, by = ab[0], ab[1], ab[3], ab[4]
    rho, lam, mu = w["model"]
    pts = (sh["srci"], sh["srcj"], sh["srctype"])
    rc = (sh["rcvi"], sh["rcvj"], sh["rcvtype"])
    r0, _ = po.elastic_forward(0, NX, NY, 1000, p.DELTAT, p.DELTAX, p.DELTAY, ax, bx, ay, by, rho, lam, mu, *pts,
                               sh["srcv"], *rc)
    src, rcv = A.ElasticSource(*pts, sh["srcv"]), A.ElasticReceiver(*rc)
    r1, _ = A.elastic_forward(p, src, rho, lam, mu, rcv, ctx=ctx)
    assert np.abs(r0).max() > 0 and np.array_equal(r1, r0)
    n2 = 100
    p2 = A.ElasticPropagatorParams(**{**p.__dict__, "NSTEP": n2})
    srcv2 = sh["srcv"][:n2] * 0 + A.Ricker(p2, 8.0, 20.0, 1e6).reshape(-1, 1)     # a wavelet that fits 100 steps
    rho_o, lam_o, mu_o = w["model_obs"]
    a2 = (0, NX, NY, n2, p.DELTAT, p.DELTAX, p.DELTAY, ax, bx, ay, by)
    # 100 steps move the wavefront ~35 cells: the receiver line of the shortened run sits 12 cells from the source
    rc = (np.linspace(NX // 2 - 30, NX // 2 + 30, 200).astype(np.int64), np.full(200, NY // 2 + 12), sh["rcvtype"])
    rcv = A.ElasticReceiver(*rc)
    obs, _ = po.elastic_forward(*a2, rho_o, lam_o, mu_o, *pts, srcv2, *rc)
    O = po.elastic_misfit_grad(*a2, rho, lam, mu, *pts, srcv2, *rc, obs)
    G = A.elastic_misfit_grad(p2, A.ElasticSource(*pts, srcv2), rho, lam, mu, rcv, obs, ctx=ctx)
    assert O["loss"] > 0 and np.array_equal(G["rcvv"], O["rcvv"]) and abs(G["loss"] - O["loss"]) <= 1e-12 * O["loss"]
    for k, ok in (("grad_rho", "grad_rho"), ("grad_lambda", "grad_lam"), ("grad_mu", "grad_mu"), ("grad_srcv", "grad_srcv")):
        assert np.abs(O[ok]).max() > 0 and relerr(G[k], O[ok]) < TOL, k
