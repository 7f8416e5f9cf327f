"""GPU parity tests for the acoustic path, through the C ABI, against the CPU oracle and the golden vectors.
Tolerances: forward wavefields / traces are BIT-IDENTICAL to the reference op (same expression order, no FMA);
misfit and gradients <= 1e-10 relative (BASELINE.json), typically 1e-14 (the gather-form adjoint sums in a
different order than the scatter-form reference)."""
import numpy as np
import pytest

from conftest import golden, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _params(A, G):
    use = [bool(x) for x in G["use"]]
    return A.AcousticPropagatorParams(PropagatorKernel=1, NX=int(G["NX"]), NY=int(G["NY"]), NSTEP=int(G["NSTEP"]), DELTAX=float(G["dx"]),
                                      DELTAY=float(G["dy"]), DELTAT=float(G["dt"]), NPOINTS_PML=int(G["npml"]),
                                      vp_ref=float(G["vp_ref"]), USE_PML_XMIN=use[0], USE_PML_XMAX=use[1],
                                      USE_PML_YMIN=use[2], USE_PML_YMAX=use[3])


@pytest.mark.parametrize("name", ["acoustic_small.npz", "acoustic_nopml_y.npz"])
def test_golden(A, ctx, name):
    G = golden(name)
    p = _params(A, G)
    src = A.AcousticSource(G["srci"], G["srcj"], G["srcv"])
    rcv = A.AcousticReceiver(G["rcvi"], G["rcvj"])
    ap = A.AcousticPropagatorSolver(p, src, G["c"], ctx=ctx)
    A.SimulatedObservation_(ap, rcv)
    assert np.array_equal(rcv.rcvv, G["rcvv"])
    u = ap.u
    assert np.array_equal(u[-1], G["u_last"]) and np.array_equal(u[p.NSTEP // 2], G["u_mid"])
    R = A.acoustic_misfit_grad(p, src, G["c"], rcv, G["obs"], ctx=ctx)
    assert np.array_equal(R["rcvv"], G["rcvv"])
    assert abs(R["loss"] - float(G["loss"])) / float(G["loss"]) < 1e-13
    assert relerr(R["grad_c"], G["grad_c"]) < TOL
    assert relerr(R["grad_srcv"], G["grad_srcv"]) < TOL


def _case(po, rng, NX, NY, NSTEP, dx, dy, dt, npml, vp_ref, nsrc=2, nrcv=40, use=(True,) * 4):
    sig, tau = po.acoustic_pml(NX, NY, dx, dy, npml=npml, vp_ref=vp_ref, use=use)
    c = vp_ref * (1 + 0.1 * rng.random((NX + 2, NY + 2)))
    srci = rng.integers(2, NX, nsrc)
    srcj = rng.integers(2, NY, nsrc)
    srcv = np.stack([po.ricker(NSTEP, 8.0 + k, 20.0 + 3 * k, 1e6) for k in range(nsrc)], 1)
    rcvi = rng.integers(1, NX + 2, nrcv)
    rcvj = rng.integers(1, NY + 2, nrcv)
    return sig, tau, c, srci, srcj, srcv, rcvi, rcvj


@pytest.mark.parametrize("shape", [(150, 700, 60), (403 - 2, 135 - 2, 80), (70, 1100, 40), (300, 64, 50)])
def test_vs_oracle_random(A, ctx, po, shape):
    """Grids wide enough to exercise the register-marching fast path, ragged widths (not multiples of 64/512), the
    PML frame on all sides, sources/receivers anywhere (including the ring), duplicates."""
    NX, NY, NSTEP = shape
    rng = np.random.default_rng(NX * 7 + NY)
    dx, dy, dt, vp = 10.0, 8.0, 1e-3, 2500.0
    sig, tau, c, srci, srcj, srcv, rcvi, rcvj = _case(po, rng, NX, NY, NSTEP, dx, dy, dt, 12, vp)
    srci[0], srcj[0] = 1, NY // 2           # a source on the ring row
    rcvi[:2], rcvj[:2] = srci[1], srcj[1]   # duplicate receivers on a source cell
    u0, r0 = po.acoustic_forward(NX, NY, NSTEP, dt, dx, dy, sig, tau, c, srci, srcj, srcv, rcvi, rcvj)
    p = A.AcousticPropagatorParams(PropagatorKernel=1, NX=NX, NY=NY, NSTEP=NSTEP, DELTAX=dx, DELTAY=dy, DELTAT=dt, vp_ref=vp)
    src, rcv = A.AcousticSource(srci, srcj, srcv), A.AcousticReceiver(rcvi, rcvj)
    plan = A.AcousticPlan(p, srci, srcj, rcvi, rcvj, ctx=ctx)
    assert plan.info()["fast_rows"] > 0
    plan.set_model(c)
    plan.set_srcv(srcv)
    plan.forward()
    assert np.array_equal(plan.rcvv(), r0)
    for s in (2, NSTEP // 2, NSTEP):
        assert np.array_equal(plan.snapshot(s), u0[s])
    obs = 0.7 * r0 + 0.02 * np.abs(r0).max() * rng.standard_normal(r0.shape)
    L0, gc0, gs0 = po.acoustic_misfit_grad(NX, NY, NSTEP, dt, dx, dy, sig, tau, c, srci, srcj, rcvi, rcvj, obs, u0)
    plan.set_obs(obs)
    plan.gradient()
    assert abs(plan.loss() - L0) / L0 < 1e-13
    assert relerr(plan.grad_c(), gc0) < TOL
    assert relerr(plan.grad_srcv(), gs0) < TOL
    plan.close()


def test_checkpointed_gradient_equals_full_history(A, ctx, po):
    """Segment checkpointing replays the forward bit-identically: gradients with a 9-snapshot window are equal to
    the full-history ones to the last bit, and both match the oracle."""
    rng = np.random.default_rng(77)
    NX, NY, NSTEP, dx, dt, vp = 90, 600, 75, 10.0, 1e-3, 2000.0
    sig, tau, c, srci, srcj, srcv, rcvi, rcvj = _case(po, rng, NX, NY, NSTEP, dx, dx, dt, 10, vp)
    p = A.AcousticPropagatorParams(PropagatorKernel=1, NX=NX, NY=NY, NSTEP=NSTEP, DELTAX=dx, DELTAY=dx, DELTAT=dt, vp_ref=vp,
                                   NPOINTS_PML=10)
    u0, r0 = po.acoustic_forward(NX, NY, NSTEP, dt, dx, dx, sig, tau, c, srci, srcj, srcv, rcvi, rcvj)
    obs = 0.5 * r0
    res = []
    for budget in (0, None):
        plan = A.AcousticPlan(p, srci, srcj, rcvi, rcvj, ctx=ctx,
                              hist_bytes_budget=0 if budget == 0 else 13 * plan_bytes)
        if budget == 0:
            plan_bytes = plan.info()["local_rows"] * plan.info()["pitch"] * 8
        plan.set_model(c); plan.set_srcv(srcv); plan.set_obs(obs)
        plan.gradient()
        info = plan.info()
        res.append((plan.loss(), plan.grad_c(), plan.grad_srcv(), plan.rcvv(), info))
        plan.close()
    assert res[0][4]["segments"] == 1 and res[1][4]["segments"] > 3 and res[1][4]["recomputed_steps"] > 0
    assert res[0][0] == res[1][0]
    for k in (1, 2, 3):
        assert np.array_equal(res[0][k], res[1][k])
    L0, gc0, gs0 = po.acoustic_misfit_grad(NX, NY, NSTEP, dt, dx, dx, sig, tau, c, srci, srcj, rcvi, rcvj, obs, u0)
    assert relerr(res[1][1], gc0) < TOL and relerr(res[1][2], gs0) < TOL


def test_mpi_convention(A, ctx, po):
    """MPIAcousticPropagatorSolver inputs: c given as c^2 on the unpadded grid, unpadded 1-based indices."""
    rng = np.random.default_rng(5)
    NX, NY, NSTEP, dx, dt = 200, 640, 50, 10.0, 0.004
    sig, tau = po.acoustic_pml(NX, NY, dx, dx, npml=12, vp_ref=1000.0, Rcoef=0.2)
    c2 = 1.0e6 * (1 + 0.2 * rng.random((NX, NY)))
    srci, srcj = np.array([NX // 5, 1]), np.array([NY // 2, 1])
    srcv = np.stack([po.ricker(NSTEP, 6.0, 15.0, 1e4)] * 2, 1)
    rcvi, rcvj = np.full(100, NX // 5), np.arange(NY // 2 - 50, NY // 2 + 50)
    c2p = np.zeros((NX + 2, NY + 2)); c2p[1:-1, 1:-1] = c2
    u0, r0 = po.acoustic_forward(NX, NY, NSTEP, dt, dx, dx, sig, tau, c2p, srci, srcj, srcv, rcvi, rcvj,
                                 mpi_convention=True)
    obs = 0.9 * r0
    L0, g0, s0 = po.acoustic_misfit_grad(NX, NY, NSTEP, dt, dx, dx, sig, tau, c2p, srci, srcj, rcvi, rcvj, obs, u0,
                                         mpi_convention=True)
    p = A.AcousticPropagatorParams(PropagatorKernel=1, NX=NX, NY=NY, NSTEP=NSTEP, DELTAX=dx, DELTAY=dx, DELTAT=dt, vp_ref=1000.0,
                                   Rcoef=0.2, mpi_convention=True)
    R = A.acoustic_misfit_grad(p, A.AcousticSource(srci, srcj, srcv), c2, A.AcousticReceiver(rcvi, rcvj), obs, ctx=ctx)
    assert np.array_equal(R["rcvv"], r0)
    assert abs(R["loss"] - L0) / L0 < 1e-13
    assert relerr(R["grad_c"], g0[1:-1, 1:-1]) < TOL and relerr(R["grad_srcv"], s0) < TOL


def test_op_level_step(A, ctx, po):
    """The reference's own op test inputs (deps/CustomOps/AcousticOneStepCpu/gradtest.jl:15-31) through the
    op-level C ABI with device pointers."""
    import torch
    G = golden("acoustic_step_gradtest.npz")
    dev = torch.device("cuda")
    ins = [torch.tensor(x, device=dev) for x in G["ins"]]
    g = [torch.tensor(x, device=dev) for x in G["g"]]
    outs = [torch.empty_like(ins[0]) for _ in range(3)]
    A.acoustic_one_step(ctx, *ins, 0.1, 0.1, 0.1, 10, 10, *outs)
    ctx.sync()
    assert np.array_equal(torch.stack(outs).cpu().numpy(), G["fwd"])
    gout = [torch.empty_like(ins[0]) for _ in range(5)]
    A.acoustic_one_step_grad(ctx, *gout, *g, ins[0], ins[4], ins[5], ins[6], 0.1, 0.1, 0.1, 10, 10)
    ctx.sync()
    assert relerr(torch.stack(gout).cpu().numpy(), G["bwd"]) < 1e-14


def test_errors_are_reported(A, ctx):
    p = A.AcousticPropagatorParams(NX=50, NY=50, NSTEP=10, PropagatorKernel=3)
    with pytest.raises(A.AdseisError):
        A.AcousticPlan(p, [5], [5], [6], [6], ctx=ctx)
    with pytest.raises(A.AdseisError):     # a PropagatorKernel=0 slab keeps two halo rows: needs at least two rows
        A.AcousticPlan(A.AcousticPropagatorParams(NX=50, NY=50, NSTEP=10, PropagatorKernel=0), [5], [5], [6], [6],
                       ctx=ctx, slab=(0, 2, 0, 1))
    # PropagatorKernel=0 slabs exist since round 2 (parity: tests/mgpu_worker.py)
    A.AcousticPlan(A.AcousticPropagatorParams(NX=50, NY=50, NSTEP=10, PropagatorKernel=0), [5], [5], [6], [6], ctx=ctx,
                   slab=(0, 2, 0, 26)).close()
    p = A.AcousticPropagatorParams(NX=50, NY=50, NSTEP=10)
    with pytest.raises(A.AdseisError):
        A.AcousticPlan(p, [500], [5], [6], [6], ctx=ctx)       # source outside the grid
    plan = A.AcousticPlan(p, [5], [5], [6], [6], ctx=ctx)
    with pytest.raises(A.AdseisError):
        plan.forward()                                          # no model yet
    with pytest.raises(A.AdseisError):
        plan.set_srcv(np.zeros((3, 1)))                         # too few rows


# ---------------------------------------------------------------------------------------------------------------
# PropagatorKernel = 0 (phi, psi driven by the new wavefield; src/Core.jl:528-549).  The reference computes this
# scheme with TF element-wise kernels, not with the C++ op, so parity is to fp64 round-off (1e-10 bar), not bit-exact.
# ---------------------------------------------------------------------------------------------------------------
def test_kernel0_golden(A, ctx):
    G = golden("acoustic_kernel0.npz")
    p = A.AcousticPropagatorParams(NX=int(G["NX"]), NY=int(G["NY"]), NSTEP=int(G["NSTEP"]), DELTAX=float(G["dx"]),
                                   DELTAY=float(G["dy"]), DELTAT=float(G["dt"]), NPOINTS_PML=int(G["npml"]),
                                   vp_ref=float(G["vp_ref"]))
    assert p.PropagatorKernel == 0                                   # the reference's default (src/Struct.jl:120)
    src, rcv = A.AcousticSource(G["srci"], G["srcj"], G["srcv"]), A.AcousticReceiver(G["rcvi"], G["rcvj"])
    R = A.acoustic_misfit_grad(p, src, G["c"], rcv, G["obs"], ctx=ctx)
    assert relerr(R["rcvv"], G["rcvv"]) < 1e-12
    assert abs(R["loss"] - float(G["loss"])) / float(G["loss"]) < 1e-12
    assert relerr(R["grad_c"], G["grad_c"]) < TOL and relerr(R["grad_srcv"], G["grad_srcv"]) < TOL


@pytest.mark.parametrize("shape,mpi", [((150, 700, 60), False), ((90, 1100, 50), True), ((64, 80, 70), False)])
def test_kernel0_vs_oracle(A, ctx, po, shape, mpi):
    """Marching box + frame, sources inside the absorbing frame / next to each other / on the ring (their injected
    part must not leak into the phi/psi-driven terms), both input conventions, checkpoint segments."""
    NX, NY, NSTEP = shape
    rng = np.random.default_rng(NX + NY)
    dx, dy, dt, vp = 10.0, 8.0, 1e-3, 2500.0
    sig, tau, c, srci, srcj, srcv, rcvi, rcvj = _case(po, rng, NX, NY, NSTEP, dx, dy, dt, 12, vp, nsrc=5)
    srci[0], srcj[0] = 1, NY // 2            # ring row
    srci[1], srcj[1] = 4, 6                  # inside the corner of the absorbing frame
    srci[2], srcj[2] = 5, 6                  # its neighbour
    srci[3], srcj[3] = NX // 2, NY - 2       # inside the y-max strip
    off = 1 if mpi else 0                    # MPI convention: 1-based into the UNPADDED grid
    cc = c * c if mpi else c
    u0, up0, r0 = po.acoustic_forward(NX, NY, NSTEP, dt, dx, dy, sig, tau, cc, srci - off, srcj - off, srcv, rcvi - off,
                                      rcvj - off, mpi_convention=mpi, kernel=0)
    obs = 0.7 * r0 + 0.02 * np.abs(r0).max() * rng.standard_normal(r0.shape)
    L0, gc0, gs0 = po.acoustic_misfit_grad(NX, NY, NSTEP, dt, dx, dy, sig, tau, cc, srci - off, srcj - off, rcvi - off,
                                           rcvj - off, obs, u0, mpi_convention=mpi, upre_hist=up0)
    p = A.AcousticPropagatorParams(NX=NX, NY=NY, NSTEP=NSTEP, DELTAX=dx, DELTAY=dy, DELTAT=dt, vp_ref=vp,
                                   PropagatorKernel=0, mpi_convention=mpi)
    model = cc[1:-1, 1:-1] if mpi else cc
    want_gc = gc0[1:-1, 1:-1] if mpi else gc0
    out = []
    for hist in (0, 9):
        plan = A.AcousticPlan(p, srci - off, srcj - off, rcvi - off, rcvj - off, ctx=ctx,
                              hist_bytes_budget=hist * (NX + 2) * ((NY + 2 + 15) // 16 * 16) * 8)
        plan.set_model(np.ascontiguousarray(model)); plan.set_srcv(srcv); plan.set_obs(obs)
        plan.gradient()
        assert relerr(plan.rcvv(), r0) < 1e-12
        assert abs(plan.loss() - L0) / L0 < 1e-12
        assert relerr(plan.grad_c(), want_gc) < TOL
        assert relerr(plan.grad_srcv(), gs0) < TOL
        out.append((plan.loss(), plan.grad_c(), plan.info()))
        plan.close()
    assert out[1][2]["segments"] > 1
    assert out[0][0] == out[1][0] and np.array_equal(out[0][1], out[1][1])


@pytest.mark.parametrize("kernel", [1, 0])
def test_two_step_kernels_equal_one_step_kernels_bitwise(A, ctx, po, monkeypatch, kernel):
    """Temporal blocking (ac_fwd2_kernel / ac_adj2_kernel + frame-only launches) is an execution schedule, not a
    different discretisation: traces, loss and both gradients are bit-identical to the one-step kernels, with sources
    and receivers inside the box, on its rim, in the frame, on tile seams, and with checkpoint segments that shift the
    pairing of the steps."""
    NX, NY, NSTEP = 150, 1300, 61
    rng = np.random.default_rng(31)
    dx, dt, vp = 10.0, 1e-3, 2500.0
    sig, tau, c, srci, srcj, srcv, rcvi, rcvj = _case(po, rng, NX, NY, NSTEP, dx, dx, dt, 10, vp, nsrc=6)
    srci[0], srcj[0] = 14, 18          # first row / column of the two-step box (npml 10: PML-free rows 12.., box 14..)
    srci[1], srcj[1] = 13, 40          # on the rim of the box (a frame cell whose value the box kernel recomputes)
    srci[2], srcj[2] = 60, 16 + 512    # first column of the second column tile
    srci[3], srcj[3] = 60, 16 + 511    # last column of the first one
    srci[4], srcj[4] = 5, 5            # deep inside the absorbing frame
    rcvi[:8] = [14, 13, 60, 60, 5, 70, 71, 72]
    rcvj[:8] = [18, 40, 16 + 512, 16 + 511, 5, 16 + 64, 16 + 63, 16 + 65]
    rcvi[8:14] = [15, 15, 16, 40, 40, 15]        # on and next to the first row / column of the box (its rim ring), duplicates
    rcvj[8:14] = [17, 18, 17, 17, 18, 17]
    srci[5], srcj[5] = 15, 17                    # a source on the corner cell of the rim ring
    p = A.AcousticPropagatorParams(PropagatorKernel=kernel, NX=NX, NY=NY, NSTEP=NSTEP, DELTAX=dx, DELTAY=dx, DELTAT=dt,
                                   vp_ref=vp, NPOINTS_PML=10)
    pitch = (NY + 2 + 15) // 16 * 16
    out = {}
    monkeypatch.setenv("ADSEIS_AC_PERSIST", "0")      # this test is about the step kernels (the whole-sweep kernel has its own)
    for tb in ("0", "1", "1s"):       # one-step kernels; pairs with frame / box launches on two streams; pairs on one stream
        monkeypatch.setenv("ADSEIS_AC_TB", tb[0])
        monkeypatch.setenv("ADSEIS_AC_TB_OVERLAP", "0" if tb == "1s" else "1")
        for slots in (0, 12):
            plan = A.AcousticPlan(p, srci, srcj, rcvi, rcvj, ctx=ctx, hist_bytes_budget=slots * (NX + 2) * pitch * 8)
            plan.set_model(c); plan.set_srcv(srcv)
            plan.forward()
            r = plan.rcvv()
            plan.set_obs(0.6 * r)
            plan.gradient()
            out[(tb, slots)] = (r, plan.loss(), plan.grad_c(), plan.grad_srcv(), plan.info())
            plan.close()
    ref = out[("0", 0)]
    assert np.abs(ref[0]).max() > 0 and np.abs(ref[2]).max() > 0
    assert out[("1", 12)][4]["segments"] > 3
    assert out[("1", 0)][4]["launches"] > ref[4]["launches"]       # three launches per pair of steps instead of two
    for key, val in out.items():
        assert np.array_equal(val[0], ref[0]) and val[1] == ref[1], key
        assert np.array_equal(val[2], ref[2]) and np.array_equal(val[3], ref[3]), key


@pytest.mark.parametrize("kernel,frame_sources", [(1, True), (0, False), (0, True)])
def test_whole_sweep_kernel_equals_step_kernels_bitwise(A, ctx, po, monkeypatch, kernel, frame_sources):
    """Small grids run a whole sweep in ONE cooperative launch (ac_fwd_persist_kernel / ac_adj_persist_kernel: resident
    CTAs, a grid barrier per step).  It is a schedule, not a discretisation: same bits as one launch per step, for both
    schemes; with PropagatorKernel=0 and sources inside the absorbing frame (c-gradient correction kernel between the
    adjoint launches) the plan falls back to the step kernels."""
    NX, NY, NSTEP = 133, 401, 90
    rng = np.random.default_rng(77)
    dx, dt, vp = 10.0, 1e-3, 2500.0
    sig, tau, c, srci, srcj, srcv, rcvi, rcvj = _case(po, rng, NX, NY, NSTEP, dx, dx, dt, 12, vp, nsrc=4)
    srci[:] = [40, 41, 70, 100]; srcj[:] = [200, 200, 30, 380]
    if frame_sources:
        srci[0], srcj[0] = 5, 6; srci[1], srcj[1] = 1, NY // 2
    rcvi[:4] = [5, 1, 40, 41]; rcvj[:4] = [6, NY // 2, 200, 200]
    p = A.AcousticPropagatorParams(PropagatorKernel=kernel, NX=NX, NY=NY, NSTEP=NSTEP, DELTAX=dx, DELTAY=dx, DELTAT=dt,
                                   vp_ref=vp, NPOINTS_PML=12)
    out = {}
    for ps in ("0", "1"):
        monkeypatch.setenv("ADSEIS_AC_PERSIST", ps)
        plan = A.AcousticPlan(p, srci, srcj, rcvi, rcvj, ctx=ctx)
        plan.set_model(c); plan.set_srcv(srcv)
        plan.forward()
        r = plan.rcvv()
        plan.set_obs(0.6 * r)
        for rep in range(2):
            plan.gradient()
        out[ps] = (r, plan.loss(), plan.grad_c(), plan.grad_srcv(), plan.snapshot(NSTEP), plan.info())
        plan.close()
    a, b = out["0"], out["1"]
    assert np.abs(a[0]).max() > 0 and np.abs(a[2]).max() > 0
    expect_persist = not (kernel == 0 and frame_sources)
    assert (b[5]["launches"] < a[5]["launches"] / 10) == expect_persist, (a[5], b[5])
    assert np.array_equal(a[0], b[0]) and a[1] == b[1] and np.array_equal(a[4], b[4])
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
