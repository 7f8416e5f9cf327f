"""Pins the CPU oracle (oracle/oracle.c): (1) against the committed golden vectors, which were produced from the
reference's own C++ op bodies / a torch-autograd restatement of its elastic graph (tests/golden/make_golden.py);
(2) against oracle/_ref directly when that library is present; (3) through the reference's own test strategy --
finite-difference convergence of the gradients (deps/CustomOps/*/gradtest.jl, examples/demo/ElasticWave_gradtest.jl).
CPU only."""
import numpy as np
import pytest

from conftest import golden, relerr


def _acoustic_inputs(po, G):
    NX, NY, NSTEP = int(G["NX"]), int(G["NY"]), int(G["NSTEP"])
    sig, tau = po.acoustic_pml(NX, NY, float(G["dx"]), float(G["dy"]), npml=int(G["npml"]), vp_ref=float(G["vp_ref"]),
                               use=tuple(bool(x) for x in G["use"]))
    return NX, NY, NSTEP, sig, tau


@pytest.mark.parametrize("name", ["acoustic_small.npz", "acoustic_nopml_y.npz"])
def test_acoustic_oracle_vs_golden(po, name):
    G = golden(name)
    NX, NY, NSTEP, sig, tau = _acoustic_inputs(po, G)
    assert np.array_equal(sig.reshape(NX + 2, NY + 2)[:, 0], G["sigx"])
    assert np.array_equal(tau.reshape(NX + 2, NY + 2)[0, :], G["tauy"])
    u, rcvv = po.acoustic_forward(NX, NY, NSTEP, float(G["dt"]), float(G["dx"]), float(G["dy"]), sig, tau, G["c"],
                                  G["srci"], G["srcj"], G["srcv"], G["rcvi"], G["rcvj"])
    # the oracle restates the reference bodies expression by expression: bit-identical
    assert np.array_equal(rcvv, G["rcvv"])
    assert np.array_equal(u[-1], G["u_last"]) and np.array_equal(u[NSTEP // 2], G["u_mid"])
    loss, gc, gs = po.acoustic_misfit_grad(NX, NY, NSTEP, float(G["dt"]), float(G["dx"]), float(G["dy"]), sig, tau,
                                           G["c"], G["srci"], G["srcj"], G["rcvi"], G["rcvj"], G["obs"], u)
    assert loss == float(G["loss"])
    assert np.array_equal(gc, G["grad_c"]) and np.array_equal(gs, G["grad_srcv"])


def test_acoustic_step_vs_golden(po):
    G = golden("acoustic_step_gradtest.npz")
    ins, g = G["ins"], G["g"]
    fwd = po.acoustic_step_fwd(*ins, 0.1, 0.1, 0.1, 10, 10)
    bwd = po.acoustic_step_bwd(*g, ins[0], ins[4], ins[5], ins[6], 0.1, 0.1, 0.1, 10, 10)
    assert np.array_equal(np.stack(fwd), G["fwd"])
    assert np.array_equal(np.stack(bwd), G["bwd"])


def test_acoustic_oracle_vs_ref_live(po):
    if not po.has_ref():
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    rng = np.random.default_rng(5)
    NX, NY, NSTEP, dx, dy, dt = 37, 29, 80, 10.0, 7.0, 8e-4
    sig, tau = po.acoustic_pml(NX, NY, dx, dy, npml=7, vp_ref=2200.0)
    c = 2200.0 * (1 + 0.1 * rng.standard_normal((NX + 2, NY + 2)))
    srci, srcj = np.array([10, 20]), np.array([12, 9])
    srcv = np.stack([po.ricker(NSTEP, 9.0, 25.0, 1e6), po.ricker(NSTEP, 7.0, 30.0, 1e6)], 1)
    rcvi, rcvj = np.arange(3, 35), np.full(32, 5)
    u1, r1 = po.acoustic_forward(NX, NY, NSTEP, dt, dx, dy, sig, tau, c, srci, srcj, srcv, rcvi, rcvj)
    u2, r2 = po.acoustic_forward(NX, NY, NSTEP, dt, dx, dy, sig, tau, c, srci, srcj, srcv, rcvi, rcvj, which="ref")
    assert np.array_equal(u1, u2) and np.array_equal(r1, r2)
    obs = 0.9 * r1
    a = po.acoustic_misfit_grad(NX, NY, NSTEP, dt, dx, dy, sig, tau, c, srci, srcj, rcvi, rcvj, obs, u1)
    b = po.acoustic_misfit_grad(NX, NY, NSTEP, dt, dx, dy, sig, tau, c, srci, srcj, rcvi, rcvj, obs, u2, which="ref")
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])


def test_acoustic_mpi_convention_matches_block_decomposed_ref(po):
    """MPIAcoustic at 2x2 blocks (reference body MpiAcousticOneStep.h + emulated halo exchange) == the oracle's
    mpi_convention path on the global grid: the decomposition invariant of examples/mpi_acoustic/verification."""
    if not po.has_ref():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(11)
    n, NSTEP, dx, dt = 16, 40, 10.0, 0.004
    NX = NY = 2 * n
    sig, tau = po.acoustic_pml(NX, NY, dx, dx, npml=6, vp_ref=1000.0, Rcoef=0.2)
    c2 = 1.0e6 * (1 + 0.2 * rng.random((NX, NY)))
    srci, srcj = np.array([NX // 5, n, n + 1]), np.array([NY // 2, n, n + 1])  # incl. block-corner cells
    srcv = np.stack([po.ricker(NSTEP, 6.0, 15.0, 1e4)] * 3, 1)
    ublk = po.ref_mpi_acoustic_forward(NX, NY, n, NSTEP, dt, dx, dx, sig, tau, c2, srci, srcj, srcv, nthreads=4)
    c2p = np.zeros((NX + 2, NY + 2))
    c2p[1:-1, 1:-1] = c2
    u, _ = po.acoustic_forward(NX, NY, NSTEP, dt, dx, dx, sig, tau, c2p, srci, srcj, srcv, [], [], mpi_convention=True)
    assert np.array_equal(u[:, 1:-1, 1:-1], ublk)


def _fd_check(f, g, x, dx, eps_list, tol):
    errs = []
    for eps in eps_list:
        fd = (f(x + eps * dx) - f(x - eps * dx)) / (2 * eps)
        an = float((g * dx).sum())
        errs.append(abs(fd - an) / abs(an))
    assert min(errs) < tol, errs


def test_acoustic_gradient_fd(po):
    rng = np.random.default_rng(3)
    NX, NY, NSTEP, dx, dt = 30, 26, 70, 10.0, 1e-3
    sig, tau = po.acoustic_pml(NX, NY, dx, dx, npml=5, vp_ref=2000.0)
    c = 2000.0 * (1 + 0.05 * rng.standard_normal((NX + 2, NY + 2)))
    srci, srcj = np.array([15]), np.array([13])
    srcv = po.ricker(NSTEP, 8.0, 20.0, 1e6).reshape(-1, 1)
    rcvi, rcvj = np.arange(2, 29), np.full(27, 3)
    u, r = po.acoustic_forward(NX, NY, NSTEP, dt, dx, dx, sig, tau, c, srci, srcj, srcv, rcvi, rcvj)
    obs = 0.8 * r + 0.01 * np.abs(r).max() * rng.standard_normal(r.shape)
    loss, gc, gs = po.acoustic_misfit_grad(NX, NY, NSTEP, dt, dx, dx, sig, tau, c, srci, srcj, rcvi, rcvj, obs, u)

    def L(c_, s_=srcv):
        return ((po.acoustic_forward(NX, NY, NSTEP, dt, dx, dx, sig, tau, c_, srci, srcj, s_, rcvi, rcvj)[1] - obs) ** 2).sum()

    _fd_check(L, gc, c, rng.standard_normal(c.shape), [1e-2, 1e-3, 1e-4], 1e-7)
    ds = rng.standard_normal(srcv.shape) * 1e5
    _fd_check(lambda s: L(c, s), gs, srcv, ds, [1e-2, 1e-3], 1e-8)


@pytest.mark.parametrize("name", ["elastic_S.npz", "elastic_M.npz"])
def test_elastic_oracle_vs_golden(po, name):
    G = golden(name)
    v, NX, NY, NSTEP = int(G["variant"]), int(G["NX"]), int(G["NY"]), int(G["NSTEP"])
    ax, bx = po.elastic_cpml_1d(NX, float(G["dx"]), float(G["dt"]), npml=int(G["npml"]), vp_ref=float(G["vp_ref"]),
                                alpha_max=float(G["alpha_max"]))
    ay, by = po.elastic_cpml_1d(NY, float(G["dy"]), float(G["dt"]), npml=int(G["npml"]), vp_ref=float(G["vp_ref"]),
                                alpha_max=float(G["alpha_max"]))
    assert np.array_equal(ax, G["ax"]) and np.array_equal(by, G["by"])
    R = po.elastic_misfit_grad(v, NX, NY, NSTEP, float(G["dt"]), float(G["dx"]), float(G["dy"]), ax, bx, ay, by,
                               G["rho"], G["lam"], G["mu"], G["srci"], G["srcj"], G["srctype"], G["srcv"], G["rcvi"],
                               G["rcvj"], G["rcvtype"], G["obs"])
    # forward: same expressions -> identical; gradients: hand-derived transpose vs autograd -> rounding only
    assert relerr(R["rcvv"], G["rcvv"]) < 1e-14
    assert abs(R["loss"] - float(G["loss"])) / float(G["loss"]) < 1e-13
    for k in ("grad_rho", "grad_lam", "grad_mu", "grad_srcv"):
        assert relerr(R[k], G[k]) < 1e-12, k


@pytest.mark.parametrize("variant", [0, 1])
def test_elastic_gradient_fd(po, variant):
    """examples/demo/ElasticWave_gradtest.jl: all five source types, finite-difference check of the whole loop."""
    rng = np.random.default_rng(21 + variant)
    NX, NY, NSTEP, h, dt = 20, 18, 12, 1.0, 1e-4
    H, W = po.elastic_dims(variant, NX, NY)
    ax, bx = po.elastic_cpml_1d(NX, h, dt, npml=4, vp_ref=3300.0, alpha_max=np.pi * 15)
    ay, by = po.elastic_cpml_1d(NY, h, dt, npml=4, vp_ref=3300.0, alpha_max=np.pi * 15)
    rho = 2800.0 * (1 + 0.1 * rng.random((H, W)))
    vp = 3000.0 * (1 + 0.1 * rng.random((H, W)))
    vs = vp / 1.732
    mu, lam = rho * vs * vs, rho * (vp * vp - 2 * vs * vs)
    srci, srcj, srctype = np.array([10, 5, 8, 12, 6]), np.array([9, 6, 8, 11, 12]), np.array([0, 1, 2, 3, 4])
    srcv = rng.standard_normal((NSTEP, 5))
    rcvi, rcvj, rcvtype = np.array([4, 8, 12, 16, 9]), np.array([5, 9, 13, 7, 9]), np.array([0, 1, 2, 3, 4])
    obs = np.zeros((5, NSTEP + 1))
    args = (variant, NX, NY, NSTEP, dt, h, h, ax, bx, ay, by)
    pts = (srci, srcj, srctype)
    R = po.elastic_misfit_grad(*args, rho, lam, mu, *pts, srcv, rcvi, rcvj, rcvtype, obs)

    def L(rho_=rho, lam_=lam, mu_=mu, s_=srcv):
        r, _ = po.elastic_forward(*args, rho_, lam_, mu_, *pts, s_, rcvi, rcvj, rcvtype)
        return ((r - obs) ** 2).sum()

    d = rng.standard_normal((H, W))
    _fd_check(lambda x: L(lam_=x), R["grad_lam"], lam, d * lam.mean(), [1e-3, 1e-4, 1e-5], 1e-6)
    _fd_check(lambda x: L(mu_=x), R["grad_mu"], mu, d * mu.mean(), [1e-3, 1e-4, 1e-5], 1e-6)
    _fd_check(lambda x: L(rho_=x), R["grad_rho"], rho, d * rho.mean(), [1e-3, 1e-4, 1e-5], 1e-6)
    _fd_check(lambda x: L(s_=x), R["grad_srcv"], srcv, rng.standard_normal(srcv.shape), [1e-1, 1e-2], 1e-8)


def test_elastic_symmetry(po):
    """Homogeneous medium + centred vy source in a square domain: vx is antisymmetric / vy symmetric under x->-x
    is broken by the staggering, so test the weaker exact invariant: the solution is independent of variant-M
    ghost values and S-variant ring values of the materials (they are never read)."""
    rng = np.random.default_rng(2)
    NX, NY, NSTEP, h, dt = 16, 16, 10, 1.0, 1e-4
    ax, bx = po.elastic_cpml_1d(NX, h, dt, npml=4, vp_ref=3300.0)
    ay, by = po.elastic_cpml_1d(NY, h, dt, npml=4, vp_ref=3300.0)
    H, W = po.elastic_dims(1, NX, NY)
    rho, lam, mu = np.full((H, W), 2800.0), np.full((H, W), 8.4e9), np.full((H, W), 8.4e9)
    pts = (np.array([8]), np.array([8]), np.array([1]))
    srcv = rng.standard_normal((NSTEP, 1))
    r1, _ = po.elastic_forward(1, NX, NY, NSTEP, dt, h, h, ax, bx, ay, by, rho, lam, mu, *pts, srcv, [5], [5], [0])
    rho2, lam2, mu2 = rho.copy(), lam.copy(), mu.copy()
    for a in (rho2, lam2, mu2):
        a[:2] = a[-2:] = 1.0
        a[:, :2] = a[:, -2:] = 1.0
    r2, _ = po.elastic_forward(1, NX, NY, NSTEP, dt, h, h, ax, bx, ay, by, rho2, lam2, mu2, *pts, srcv, [5], [5], [0])
    assert np.array_equal(r1, r2)


def test_block_decomposed_ref_gradient_matches_oracle(po):
    """Gradient through the reference's per-block backward body + transposed halo exchange (2x2 blocks, threads) ==
    the oracle's reverse sweep on the global grid (decomposed == undecomposed,
    examples/mpi_elastic/verification/verify_backward.jl's invariant, for the acoustic custom op)."""
    if not po.has_ref():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(12)
    n, NSTEP, dx, dt = 20, 45, 10.0, 0.004
    NX, NY = 2 * n, 3 * n
    sig, tau = po.acoustic_pml(NX, NY, dx, dx, npml=6, vp_ref=1000.0, Rcoef=0.2)
    c2 = 1.0e6 * (1 + 0.2 * rng.random((NX, NY)))
    srci, srcj = np.array([NX // 5, n, n + 1]), np.array([NY // 2, n, n + 1])
    srcv = np.stack([po.ricker(NSTEP, 6.0, 15.0, 1e4)] * 3, 1)
    rcvi, rcvj = np.array([n, n + 1, 5, 2 * n, 7]), np.array([n, n + 1, 2 * n, 2 * n + 1, 31])
    ublk = po.ref_mpi_acoustic_forward(NX, NY, n, NSTEP, dt, dx, dx, sig, tau, c2, srci, srcj, srcv, nthreads=3)
    obs = 0.5 * ublk[:, rcvi - 1, rcvj - 1]
    Lb, gb, sb = po.ref_mpi_acoustic_gradient(NX, NY, n, NSTEP, dt, dx, dx, sig, tau, c2, srci, srcj, rcvi, rcvj, obs,
                                              ublk, nthreads=3)
    c2p = np.zeros((NX + 2, NY + 2))
    c2p[1:-1, 1:-1] = c2
    u, r = po.acoustic_forward(NX, NY, NSTEP, dt, dx, dx, sig, tau, c2p, srci, srcj, srcv, rcvi, rcvj,
                               mpi_convention=True)
    assert np.array_equal(u[:, 1:-1, 1:-1], ublk)
    L, g, s = po.acoustic_misfit_grad(NX, NY, NSTEP, dt, dx, dx, sig, tau, c2p, srci, srcj, rcvi, rcvj, obs, u,
                                      mpi_convention=True)
    assert abs(L - Lb) / L < 1e-14
    assert relerr(gb, g[1:-1, 1:-1]) < 1e-13 and relerr(sb, s) < 1e-13


# ---------------------------------------------------------------------------------------------------------------
# PropagatorKernel = 0 (src/Core.jl:528-549): pinned by a torch-autograd restatement of the reference's graph
# ---------------------------------------------------------------------------------------------------------------
def _k0_run(po, G):
    NX, NY, NSTEP = int(G["NX"]), int(G["NY"]), int(G["NSTEP"])
    dx, dy, dt = float(G["dx"]), float(G["dy"]), float(G["dt"])
    sig, tau = po.acoustic_pml(NX, NY, dx, dy, npml=int(G["npml"]), vp_ref=float(G["vp_ref"]))
    u, up, r = po.acoustic_forward(NX, NY, NSTEP, dt, dx, dy, sig, tau, G["c"], G["srci"], G["srcj"], G["srcv"],
                                   G["rcvi"], G["rcvj"], kernel=0)
    L, gc, gs = po.acoustic_misfit_grad(NX, NY, NSTEP, dt, dx, dy, sig, tau, G["c"], G["srci"], G["srcj"], G["rcvi"],
                                        G["rcvj"], G["obs"], u, upre_hist=up)
    return r, L, gc, gs


def test_acoustic_kernel0_oracle_vs_golden(po):
    G = golden("acoustic_kernel0.npz")
    r, L, gc, gs = _k0_run(po, G)
    assert relerr(r, G["rcvv"]) < 1e-13
    assert abs(L - float(G["loss"])) / float(G["loss"]) < 1e-13
    assert relerr(gc, G["grad_c"]) < 1e-12 and relerr(gs, G["grad_srcv"]) < 1e-12


@pytest.mark.parametrize("kernel", [0, 2])
def test_acoustic_torch_restatement_live(po, kernel):
    """kernel=2 (acoustic_one_step_customop_ref, Core.jl:504-525) ties the torch restatement to the C++ op bodies'
    oracle; kernel=0 ties the hand-derived scheme-0 reverse sweep to autograd of the same restatement."""
    import torch
    from oracle import torch_acoustic as ta
    rng = np.random.default_rng(11)
    NX, NY, NSTEP, dx, dy, dt = 22, 27, 40, 10.0, 8.0, 1e-3
    sig, tau = po.acoustic_pml(NX, NY, dx, dy, npml=5, vp_ref=2500.0)
    c = 2500 * (1 + 0.1 * rng.random((NX + 2, NY + 2)))
    srci, srcj = np.array([11, 3, 4, 20]), np.array([13, 5, 5, 26])
    srcv = np.stack([po.ricker(NSTEP, 6.0 + k, 10.0 + k, 1e6) for k in range(4)], 1)
    rcvi, rcvj = rng.integers(1, NX + 3, 20), rng.integers(1, NY + 3, 20)
    if kernel == 0:
        u, up, r = po.acoustic_forward(NX, NY, NSTEP, dt, dx, dy, sig, tau, c, srci, srcj, srcv, rcvi, rcvj, kernel=0)
    else:
        u, r = po.acoustic_forward(NX, NY, NSTEP, dt, dx, dy, sig, tau, c, srci, srcj, srcv, rcvi, rcvj)
        up = None
    obs = 0.7 * r + 0.01 * np.abs(r).max() * rng.standard_normal(r.shape)
    L0, gc, gs = po.acoustic_misfit_grad(NX, NY, NSTEP, dt, dx, dy, sig, tau, c, srci, srcj, rcvi, rcvj, obs, u,
                                         upre_hist=up)
    ct, st = torch.tensor(c.reshape(-1), requires_grad=True), torch.tensor(srcv, requires_grad=True)
    L, rt = ta.acoustic_loss(kernel, NX, NY, NSTEP, dt, dx, dy, sig, tau, ct, srci, srcj, st, rcvi, rcvj, obs)
    L.backward()
    assert relerr(rt.detach().numpy(), r) < 1e-13 and abs(float(L.detach()) - L0) / L0 < 1e-13
    assert relerr(ct.grad.numpy().reshape(NX + 2, NY + 2), gc) < 1e-12
    assert relerr(st.grad.numpy()[:NSTEP], gs) < 1e-12


def test_acoustic_kernel0_differs_from_kernel1(po):
    """The two schemes are different discretisations of the PML memory variables: equal outside PML influence only."""
    G = golden("acoustic_kernel0.npz")
    NX, NY, NSTEP = int(G["NX"]), int(G["NY"]), int(G["NSTEP"])
    dx, dy, dt = float(G["dx"]), float(G["dy"]), float(G["dt"])
    sig, tau = po.acoustic_pml(NX, NY, dx, dy, npml=int(G["npml"]), vp_ref=float(G["vp_ref"]))
    _, r1 = po.acoustic_forward(NX, NY, NSTEP, dt, dx, dy, sig, tau, G["c"], G["srci"], G["srcj"], G["srcv"], G["rcvi"],
                                G["rcvj"])
    d = relerr(r1, G["rcvv"])
    assert 1e-8 < d < 0.5


def test_acoustic_kernel0_block_decomposed_equals_global(po):
    """The reference's distributed invariant (examples/mpi_acoustic/verification/verify_forward.jl:78-95) for
    PropagatorKernel=0: a NumPy restatement of MPIAcoustic.jl's block-decomposed one_step (2x3 blocks, halo exchanges
    emulated) == the oracle's global-grid scheme-0 loop under the MPI input convention."""
    from oracle import np_mpi_acoustic as nm
    rng = np.random.default_rng(21)
    n, NSTEP, dx, dy, dt = 12, 45, 10.0, 9.0, 0.004
    NX, NY = 2 * n, 3 * n
    sig, tau = po.acoustic_pml(NX, NY, dx, dy, npml=5, vp_ref=1000.0, Rcoef=0.2)
    c2 = 1.0e6 * (1 + 0.2 * rng.random((NX, NY)))
    srci, srcj = np.array([NX // 5, n, n + 1, 2]), np.array([NY // 2, n, n + 1, 3])   # block corners, inside the PML
    srcv = np.stack([po.ricker(NSTEP, 6.0 + k, 12.0 + k, 1e4) for k in range(4)], 1)
    ublk = nm.mpi_acoustic_forward_k0(NX, NY, n, NSTEP, dt, dx, dy, sig, tau, c2, srci, srcj, srcv)
    c2p = np.zeros((NX + 2, NY + 2))
    c2p[1:-1, 1:-1] = c2
    u, _, _ = po.acoustic_forward(NX, NY, NSTEP, dt, dx, dy, sig, tau, c2p, srci, srcj, srcv, [], [],
                                  mpi_convention=True, kernel=0)
    assert np.abs(ublk).max() > 0 and relerr(u[:, 1:-1, 1:-1], ublk) < 1e-13
    # and the two schemes really differ on this case (the test would not notice a scheme-1 oracle otherwise)
    u1, _ = po.acoustic_forward(NX, NY, NSTEP, dt, dx, dy, sig, tau, c2p, srci, srcj, srcv, [], [], mpi_convention=True)
    assert relerr(u1[:, 1:-1, 1:-1], ublk) > 1e-8


def test_acoustic_kernel0_gradient_fd(po):
    """The reference's own gradient test strategy (finite-difference convergence, deps/CustomOps/*/gradtest.jl) applied
    to the hand-derived PropagatorKernel=0 reverse sweep: sources inside the absorbing frame, so the injected-source
    handling of the phibar / psibar terms is exercised."""
    rng = np.random.default_rng(31)
    NX, NY, NSTEP, dx, dt = 24, 30, 50, 10.0, 1e-3
    sig, tau = po.acoustic_pml(NX, NY, dx, dx, npml=6, vp_ref=2500.0)
    c = 2500.0 * (1 + 0.1 * rng.random((NX + 2, NY + 2)))
    srci, srcj = np.array([12, 3, 4]), np.array([15, 4, 4])
    srcv = np.stack([po.ricker(NSTEP, 6.0 + k, 10.0 + k, 1e6) for k in range(3)], 1)
    rcvi, rcvj = rng.integers(2, NX + 1, 12), rng.integers(2, NY + 1, 12)

    def fwd(c_, s_):
        return po.acoustic_forward(NX, NY, NSTEP, dt, dx, dx, sig, tau, c_, srci, srcj, s_, rcvi, rcvj, kernel=0)

    u, up, r = fwd(c, srcv)
    obs = 0.8 * r + 0.01 * np.abs(r).max() * rng.standard_normal(r.shape)
    loss, gc, gs = po.acoustic_misfit_grad(NX, NY, NSTEP, dt, dx, dx, sig, tau, c, srci, srcj, rcvi, rcvj, obs, u,
                                           upre_hist=up)
    L = lambda c_, s_=srcv: ((fwd(c_, s_)[2] - obs) ** 2).sum()
    assert abs(L(c) - loss) <= 1e-12 * loss
    _fd_check(L, gc, c, rng.standard_normal(c.shape), [1e-2, 1e-3, 1e-4], 1e-7)
    _fd_check(lambda s: L(c, s), gs, srcv, rng.standard_normal(srcv.shape) * 1e5, [1e-2, 1e-3], 1e-8)
