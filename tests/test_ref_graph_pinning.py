"""Pins oracle/oracle.c -- elastic variants S and M and the acoustic PropagatorKernel=0 scheme -- against the
REFERENCE'S OWN op bodies.  The reference has no hand-written adjoint for these paths: they are TensorFlow graphs over
its gather / scatter_add / scatter_nd / add_source / get_receive custom ops (src/Core.jl:31-228, 528-620;
src/MPIElastic.jl:374-645; src/MPIAcoustic.jl:212-404), differentiated by tf.gradients.  oracle/ref_graph.inc records
those graphs statement by statement, executes every custom-op node with the reference's #included forward body and
differentiates with the reference's backward bodies (oracle/_ref/libadseis_ref.so).  Bars: forward bit-identical
(same IEEE expressions), gradients <= 1e-13 relative (summation order only).  Every test is collected twice so that it
also runs on the GPU box's `-m gpu` record (it needs no GPU; oracle/_ref travels prebuilt)."""
import numpy as np
import pytest

from conftest import golden, on_both_records, relerr


@pytest.fixture(autouse=True)
def _need_ref(po):
    if not po.has_ref():
        pytest.skip("oracle/_ref not built (no /root/reference on this machine and no prebuilt library)")


def _elastic_case(po, variant, NX, NY, NSTEP, seed, npml=5):
    rng = np.random.default_rng(seed)
    dx, dy, dt = 1.0, 1.25, 1e-4
    H, W = po.elastic_dims(variant, NX, NY)
    kw = dict(npml=npml, vp_ref=3300.0, alpha_max=np.pi * 15)
    ax, bx = po.elastic_cpml_1d(NX, dx, dt, **kw)
    ay, by = po.elastic_cpml_1d(NY, dy, dt, **kw)
    vp = 3000.0 * (1 + 0.1 * rng.random((H, W)))
    vs = vp / 1.732 * (1 + 0.05 * rng.random((H, W)))
    rho = 2800.0 * (1 + 0.1 * rng.random((H, W)))
    mu, lam = rho * vs * vs, rho * (vp * vp - 2 * vs * vs)
    # all five source / receiver types, duplicates on one cell, points next to the region edges where fw1..fw4's
    # update regions differ (Core.jl:100-107, 132-139, 162-169, 190-198)
    srci = np.array([NX // 2, 2, NX // 2, 7, 9, NX - 1, NX // 2, 1, NX])
    srcj = np.array([NY // 2, 2, NY // 2, 8, 3, NY - 1, NY // 2, 1, NY])
    srctype = np.array([0, 1, 2, 3, 4, 2, 0, 4, 1])
    if variant == 0:   # padded indices: 1-based into (NX+2) x (NY+2) -- also exercise the ring rows
        srci[7], srcj[7] = 1, 5
        srci[8], srcj[8] = NX + 2, NY + 2
    srcv = rng.standard_normal((NSTEP, len(srci))) * 1e3
    rcvi = np.concatenate([rng.integers(1, NX + 1, 12), [NX // 2, NX // 2, 1, NX]])
    rcvj = np.concatenate([rng.integers(1, NY + 1, 12), [NY // 2, NY // 2, 1, NY]])
    rcvtype = np.concatenate([rng.integers(0, 5, 12), [2, 2, 0, 3]])
    return (variant, NX, NY, NSTEP, dt, dx, dy, ax, bx, ay, by, rho, lam, mu, srci, srcj, srctype, srcv, rcvi, rcvj,
            rcvtype), rng


@on_both_records()
@pytest.mark.parametrize("variant,NX,NY,block", [(0, 26, 22, None), (0, 19, 33, None), (1, 26, 22, None),
                                                 (1, 24, 20, (12, 10)), (1, 24, 24, (8, 8)), (1, 30, 18, (10, 18))])
def test_elastic_oracle_equals_reference_op_graph(po, record, variant, NX, NY, block):
    """fw1..fw4 region bounds, half / integer CPML index choice, makevector zero fill, add_source order, get_receive
    layout: oracle.c == the graph over the reference's own ops; variant M also block-decomposed with emulated
    mpi_halo_exchange2 (decomposed == undecomposed, examples/mpi_elastic/verification/verify_backward.jl:21-30)."""
    args, rng = _elastic_case(po, variant, NX, NY, 24, 100 * variant + NX)
    r0, h0 = po.elastic_forward(*args, want_hist=True)
    obs = r0 * (1 + 0.2 * rng.standard_normal(r0.shape)) + 0.05 * np.abs(r0).max() * rng.standard_normal(r0.shape)
    O = po.elastic_misfit_grad(*args, obs)
    R = po.ref_elastic(*args, obs, want_hist=True, block=block)
    assert np.abs(r0).max() > 0
    assert np.array_equal(R["rcvv"], r0) and np.array_equal(R["hist"], h0)         # bit for bit
    assert abs(R["loss"] - O["loss"]) <= 1e-14 * O["loss"]
    for k in ("grad_rho", "grad_lam", "grad_mu", "grad_srcv"):
        assert np.abs(O[k]).max() > 0 and relerr(R[k], O[k]) < 1e-13, k


@on_both_records()
@pytest.mark.parametrize("name", ["elastic_S.npz", "elastic_M.npz"])
def test_elastic_golden_is_the_reference_op_graph(po, record, name):
    """The committed elastic golden vectors are outputs of the reference-op graph (tests/golden/make_golden.py)."""
    G = golden(name)
    v, NX, NY, NSTEP = int(G["variant"]), int(G["NX"]), int(G["NY"]), int(G["NSTEP"])
    R = po.ref_elastic(v, NX, NY, NSTEP, float(G["dt"]), float(G["dx"]), float(G["dy"]), G["ax"], G["bx"], G["ay"],
                       G["by"], G["rho"], G["lam"], G["mu"], G["srci"], G["srcj"], G["srctype"], G["srcv"], G["rcvi"],
                       G["rcvj"], G["rcvtype"], G["obs"])
    assert np.array_equal(R["rcvv"], G["rcvv"]) and R["loss"] == float(G["loss"])
    for k in ("grad_rho", "grad_lam", "grad_mu", "grad_srcv"):
        assert np.array_equal(R[k], G[k]), k


def _acoustic_case(po, NX, NY, NSTEP, seed):
    rng = np.random.default_rng(seed)
    dx, dy, dt, npml, vp = 10.0, 8.0, 1e-3, 6, 2500.0
    sig, tau = po.acoustic_pml(NX, NY, dx, dy, npml=npml, vp_ref=vp)
    c = vp * (1 + 0.1 * rng.random((NX + 2, NY + 2)))
    # interior, inside the absorbing frame on neighbouring cells, on the ring, in a corner
    srci = np.array([NX // 2, 3, 4, 1, NX - 1], dtype=np.int64)
    srcj = np.array([NY // 2, 5, 5, NY // 3, NY], dtype=np.int64)
    srcv = np.stack([po.ricker(NSTEP, 6.0 + k, 10.0 + k, 1e6) for k in range(5)], 1)
    rcvi, rcvj = rng.integers(1, NX + 3, 24), rng.integers(1, NY + 3, 24)
    return (NX, NY, NSTEP, dt, dx, dy, sig, tau, c, srci, srcj, srcv, rcvi, rcvj), rng


@on_both_records()
@pytest.mark.parametrize("kernel", [0, 2])
def test_acoustic_oracle_equals_reference_op_graph(po, record, kernel):
    """kernel 0: `one_step` (Core.jl:528-549) -- oracle.c's hand-derived scheme-0 sweep vs the gather / scatter_nd
    graph, bit-identical forward.  kernel 2: `acoustic_one_step_customop_ref` (Core.jl:504-525), the op-free twin of
    the C++ custom op, vs the oracle's restatement of that op (AcousticOneStepCpu.h): equal to round-off only (the
    C++ body groups the products differently) -- the reference's own PropagatorKernel=1 == 2 equivalence."""
    args, rng = _acoustic_case(po, 30, 37, 60, 17)
    NX, NY, NSTEP, dt, dx, dy, sig, tau, c, srci, srcj, srcv, rcvi, rcvj = args
    if kernel == 0:
        u, up, r0 = po.acoustic_forward(*args, kernel=0)
    else:
        (u, r0), up = po.acoustic_forward(*args), None
    obs = 0.7 * r0 + 0.01 * np.abs(r0).max() * rng.standard_normal(r0.shape)
    L, gc, gs = po.acoustic_misfit_grad(NX, NY, NSTEP, dt, dx, dy, sig, tau, c, srci, srcj, rcvi, rcvj, obs, u,
                                        upre_hist=up)
    R = po.ref_acoustic_graph(kernel, *args, obs, want_hist=True)
    if kernel == 0:
        assert np.array_equal(R["u"], u) and np.array_equal(R["rcvv"], r0) and R["loss"] == L
        tol = 1e-13
    else:
        assert relerr(R["u"], u) < 1e-13 and relerr(R["rcvv"], r0) < 1e-12
        tol = 1e-12
    assert relerr(R["grad_c"], gc) < tol and relerr(R["grad_srcv"], gs) < tol


@on_both_records()
@pytest.mark.parametrize("kernel,block", [(0, (36, 36)), (0, (12, 12)), (0, (18, 12)), (2, (12, 12))])
def test_mpi_acoustic_blocks_equal_global_oracle(po, record, kernel, block):
    """MPIAcoustic.jl:212-246 (scheme 0: the extra exchange of the NEW wavefield at :236) and :297-326 on 1x1, 3x3
    and 2x3 blocks with emulated mpi_halo_exchange == the oracle's global grid under the MPI input convention
    (test/verify_forward.jl:32-90's invariant), forward AND gradient."""
    rng = np.random.default_rng(3)
    n, NSTEP, dx, dt = 12, 50, 10.0, 0.004
    NX = NY = 3 * n
    sig, tau = po.acoustic_pml(NX, NY, dx, dx, npml=6, vp_ref=1000.0, Rcoef=0.2)
    c2 = 1.0e6 * (1 + 0.2 * rng.random((NX, NY)))
    srci, srcj = np.array([NX // 5, n, n + 1, 2]), np.array([NY // 2, n, n + 1, 3])   # block corners, inside the PML
    srcv = np.stack([po.ricker(NSTEP, 6.0 + k, 15.0, 1e4) for k in range(4)], 1)
    rcvi, rcvj = rng.integers(1, NX + 1, 20), rng.integers(1, NY + 1, 20)
    c2p = np.zeros((NX + 2, NY + 2))
    c2p[1:-1, 1:-1] = c2
    a = (NX, NY, NSTEP, dt, dx, dx, sig, tau, c2p, srci, srcj, srcv, rcvi, rcvj)
    if kernel == 0:
        u, up, r0 = po.acoustic_forward(*a, mpi_convention=True, kernel=0)
    else:
        (u, r0), up = po.acoustic_forward(*a, mpi_convention=True), None
    obs = 0.8 * r0
    L, gc, gs = po.acoustic_misfit_grad(NX, NY, NSTEP, dt, dx, dx, sig, tau, c2p, srci, srcj, rcvi, rcvj, obs, u,
                                        mpi_convention=True, upre_hist=up)
    R = po.ref_mpi_acoustic_graph(kernel, NX, NY, block, NSTEP, dt, dx, dx, sig, tau, c2, srci, srcj, srcv, rcvi, rcvj,
                                  obs, want_hist=True)
    if kernel == 0:
        assert np.array_equal(R["u"], u[:, 1:-1, 1:-1])
    else:
        assert relerr(R["u"], u[:, 1:-1, 1:-1]) < 1e-13
    assert abs(R["loss"] - L) <= 1e-13 * L
    assert relerr(R["grad_c2"], gc[1:-1, 1:-1]) < 1e-12 and relerr(R["grad_srcv"], gs) < 1e-12


@on_both_records()
def test_acoustic_kernel0_golden_is_the_reference_op_graph(po, record):
    G = golden("acoustic_kernel0.npz")
    NX, NY, NSTEP = int(G["NX"]), int(G["NY"]), int(G["NSTEP"])
    dx, dy, dt = float(G["dx"]), float(G["dy"]), float(G["dt"])
    sig, tau = po.acoustic_pml(NX, NY, dx, dy, npml=int(G["npml"]), vp_ref=float(G["vp_ref"]))
    R = po.ref_acoustic_graph(0, NX, NY, NSTEP, dt, dx, dy, sig, tau, G["c"], G["srci"], G["srcj"], G["srcv"],
                              G["rcvi"], G["rcvj"], G["obs"])
    assert np.array_equal(R["rcvv"], G["rcvv"]) and R["loss"] == float(G["loss"])
    assert np.array_equal(R["grad_c"], G["grad_c"]) and np.array_equal(R["grad_srcv"], G["grad_srcv"])


@on_both_records()
def test_oracle_equals_reference_bodies_acoustic(po, record):
    """Scheme 1 (the C++ custom op): oracle.c == AcousticOneStepCpu.h bodies driven in Core.jl:562-620 order, bit for
    bit, forward and gradient -- the round-1 pin, repeated here so that it is on the GPU box's record too."""
    args, rng = _acoustic_case(po, 37, 29, 80, 5)
    NX, NY, NSTEP, dt, dx, dy, sig, tau, c, srci, srcj, srcv, rcvi, rcvj = args
    rcvi, rcvj = np.clip(rcvi, 1, NX + 2), np.clip(rcvj, 1, NY + 2)
    u1, r1 = po.acoustic_forward(NX, NY, NSTEP, dt, dx, dy, sig, tau, c, srci, srcj, srcv, rcvi, rcvj)
    u2, r2 = po.acoustic_forward(NX, NY, NSTEP, dt, dx, dy, sig, tau, c, srci, srcj, srcv, rcvi, rcvj, which="ref")
    assert np.array_equal(u1, u2) and np.array_equal(r1, r2)
    obs = 0.9 * r1
    a = po.acoustic_misfit_grad(NX, NY, NSTEP, dt, dx, dy, sig, tau, c, srci, srcj, rcvi, rcvj, obs, u1)
    b = po.acoustic_misfit_grad(NX, NY, NSTEP, dt, dx, dy, sig, tau, c, srci, srcj, rcvi, rcvj, obs, u2, which="ref")
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
