"""Size-independent properties at BASELINE.json's full grid sizes (the CPU oracle cannot run these in seconds):
  * linearity in the source: scaling srcv by a power of two scales every trace EXACTLY (bit for bit);
  * the adjoint identity: with zero observed data the misfit is L = |J s|^2, so <dL/ds, s> = 2 L -- the
    source-gradient sweep is the exact transpose of the forward sweep;
  * directional finite difference of the model gradient (the reference's own test strategy,
    deps/CustomOps/*/gradtest.jl) -- central differences agree with <g, delta>;
  * checkpoint-segmented reverse sweep == fully resident tape (bit-identical), at full width.
Time steps are shortened (the properties do not depend on NSTEP); widths/heights are the full C4 / C5 / C3 grids."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _layered(shape, v0, v1, rng):
    n, m = shape
    z = np.linspace(0, 1, m)[None, :]
    return (v0 + (v1 - v0) * np.floor(z * 4) / 4) * (1 + 0.02 * rng.random(shape))


def _exactly_scaled(r, r1, k):
    """r == k r1 bit for bit; the numerical precursor ahead of the wavefront underflows into the denormal range,
    where scaling is not exact -- there only the magnitude is checked."""
    big = np.abs(r1) > 1e-280
    return np.array_equal(r[big], k * r1[big]) and np.all(np.abs(r[~big] - k * r1[~big]) < 1e-270)


def _acoustic_plan(A, ctx, NX, NY, NSTEP, nrcv, hist=0):
    p = A.AcousticPropagatorParams(PropagatorKernel=1, NX=NX, NY=NY, NSTEP=NSTEP, DELTAX=10.0, DELTAY=10.0, DELTAT=1e-3, vp_ref=2500.0)
    srci, srcj = np.array([NX // 2, NX // 3]), np.array([NY // 2, 40])
    rcvi = np.linspace(20, NX - 20, nrcv).astype(np.int64)
    rcvj = np.full(nrcv, NY // 2 + 30)
    plan = A.AcousticPlan(p, srci, srcj, rcvi, rcvj, ctx=ctx, hist_bytes_budget=hist)
    srcv = np.stack([A.Ricker(p, 30.0, 40.0, 1e6), A.Ricker(p, 25.0, 50.0, 5e5)], 1)[:NSTEP]
    return p, plan, srcv


@pytest.mark.parametrize("grid", [(4096, 4096, 120), (2000, 1000, 150)])   # C4 and the C3 Marmousi-shaped grid
def test_acoustic_fullsize_properties(A, ctx, grid):
    NX, NY, NSTEP = grid
    rng = np.random.default_rng(1)
    p, plan, srcv = _acoustic_plan(A, ctx, NX, NY, NSTEP, 256)
    c = _layered(plan.model_shape, 1500.0, 3500.0, rng)
    plan.set_model(c); plan.set_srcv(srcv)
    plan.forward()
    r1 = plan.rcvv()
    assert np.abs(r1).max() > 0
    plan.set_srcv(4.0 * srcv)
    plan.forward()
    assert _exactly_scaled(plan.rcvv(), r1, 4.0)                              # exact linearity
    # adjoint identity
    plan.set_srcv(srcv); plan.set_obs(np.zeros_like(r1))
    plan.gradient()
    L, gs, gc = plan.loss(), plan.grad_srcv(), plan.grad_c()
    assert abs(L - float((r1 * r1).sum())) / L < 1e-12
    assert abs(float((gs * srcv).sum()) - 2 * L) / (2 * L) < 1e-11
    # directional FD of dL/dc
    d = rng.standard_normal(c.shape) * c.mean()
    eps = 1e-6
    Ls = []
    for sgn in (+1, -1):
        plan.set_model(c + sgn * eps * d)
        plan.gradient()
        Ls.append(plan.loss())
    fd, an = (Ls[0] - Ls[1]) / (2 * eps), float((gc * d).sum())
    assert abs(fd - an) / abs(an) < 1e-6, (fd, an)
    plan.close()


def test_acoustic_fullwidth_segmented_equals_resident(A, ctx):
    NX, NY, NSTEP = 4096, 4096, 60
    rng = np.random.default_rng(2)
    out = []
    for hist in (0, 14 * (NX + 2) * 4112 * 8):          # all snapshots resident / a 14-snapshot window
        p, plan, srcv = _acoustic_plan(A, ctx, NX, NY, NSTEP, 64, hist=hist)
        if not out:
            c = _layered(plan.model_shape, 1500.0, 3500.0, rng)
        plan.set_model(c); plan.set_srcv(srcv); plan.set_obs(np.zeros((NSTEP + 1, 64)))
        plan.gradient()
        out.append((plan.loss(), plan.grad_c(), plan.grad_srcv(), plan.info()))
        plan.close()
    assert out[0][3]["segments"] == 1 and out[1][3]["segments"] > 1 and out[1][3]["recomputed_steps"] > 0
    assert out[0][0] == out[1][0] and np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2], out[1][2])


@pytest.mark.parametrize("variant", [1, 0])          # C5: 2000^2, variant M (mpi_elastic analogue) and variant S
def test_elastic_fullsize_properties(A, ctx, variant):
    NX = NY = 2000
    NSTEP = 80
    rng = np.random.default_rng(3)
    p = A.ElasticPropagatorParams(NX=NX, NY=NY, NSTEP=NSTEP, DELTAX=10.0, DELTAY=10.0, DELTAT=1e-3, vp_ref=3000.0,
                                  f0=10.0, variant=variant)
    shape = p.model_shape()
    vp = _layered(shape, 2500.0, 3500.0, rng)
    vs, rho = vp / 1.732, np.full(shape, 2500.0) * (1 + 0.02 * rng.random(shape))
    lam, mu = rho * (vp * vp - 2 * vs * vs), rho * vs * vs
    nsrc = 4
    srci, srcj = np.array([1000, 700, 1200, 1000]), np.array([1000, 900, 1100, 1000])
    srctype = np.array([2, 3, 0, 4])
    srcv = np.stack([A.Ricker(p, 20.0 + k, 30.0, 1e6) for k in range(nsrc)], 1)[:NSTEP]
    nrcv = 128
    rcvi = np.linspace(900, 1100, nrcv).astype(np.int64)
    rcvj = np.full(nrcv, 1010)
    rcvtype = np.arange(nrcv) % 5
    plan = A.ElasticPlan(p, srci, srcj, srctype, rcvi, rcvj, rcvtype, ctx=ctx)
    plan.set_model(rho, lam, mu); plan.set_srcv(srcv)
    plan.forward()
    r1 = plan.rcvv()
    assert np.abs(r1).max() > 0
    plan.set_srcv(2.0 * srcv)
    plan.forward()
    assert _exactly_scaled(plan.rcvv(), r1, 2.0)
    plan.set_srcv(srcv); plan.set_obs(np.zeros_like(r1))
    plan.gradient(False)                                  # source-time-function gradient: no tape at all
    L, gs = plan.loss(), plan.grad_srcv()
    assert abs(float((gs * srcv).sum()) - 2 * L) / (2 * L) < 1e-11
    plan.gradient(True)
    gs2, gmu = plan.grad_srcv(), plan.grad_mu()
    assert np.abs(gs2 - gs).max() <= 1e-12 * np.abs(gs).max()
    d = rng.standard_normal(shape) * mu.mean()
    eps = 1e-6
    Ls = []
    for sgn in (+1, -1):
        plan.set_model(rho, lam, mu + sgn * eps * d)
        plan.gradient(False)
        Ls.append(plan.loss())
    fd, an = (Ls[0] - Ls[1]) / (2 * eps), float((gmu * d).sum())
    assert abs(fd - an) / abs(an) < 1e-5, (fd, an)
    plan.close()


def test_acoustic_forward_is_deterministic_at_full_size(A, ctx):
    """Repeat runs of one forward sweep reproduce traces and the last snapshot bit for bit.  Guards the TMA ring of
    the marching CTAs: without the generic->async proxy fence before a stage refill (ring_refill_fence, common.cuh) a
    refill could overtake a consumer's pending loads at 2 CTAs/SM -- 10 % to 100 % of such runs differed
    (profiles/r02_ring_race.md); with it 0 of 249."""
    NX, NY, NSTEP = 4096, 4096, 120
    rng = np.random.default_rng(1)
    p, plan, srcv = _acoustic_plan(A, ctx, NX, NY, NSTEP, 256)
    c = _layered(plan.model_shape, 1500.0, 3500.0, rng)
    plan.set_model(c); plan.set_srcv(srcv)
    plan.forward()
    r0, u0 = plan.rcvv(), plan.snapshot(NSTEP)
    for _ in range(8):
        plan.forward()
        assert np.array_equal(plan.rcvv(), r0) and np.array_equal(plan.snapshot(NSTEP), u0)
    plan.close()


def test_acoustic_gradient_is_deterministic_at_full_size(A, ctx):
    """The reverse sweep (ac_adj_kernel's five-plane ring, 2 CTAs/SM) reproduces loss and both gradients bit for bit."""
    NX, NY, NSTEP = 4096, 4096, 80
    rng = np.random.default_rng(4)
    p, plan, srcv = _acoustic_plan(A, ctx, NX, NY, NSTEP, 256)
    c = _layered(plan.model_shape, 1500.0, 3500.0, rng)
    plan.set_model(c); plan.set_srcv(srcv)
    plan.forward()
    plan.set_obs(0.7 * plan.rcvv())
    plan.gradient()
    L0, g0, s0 = plan.loss(), plan.grad_c(), plan.grad_srcv()
    assert np.abs(g0).max() > 0
    for _ in range(6):
        plan.gradient()
        assert plan.loss() == L0 and np.array_equal(plan.grad_c(), g0) and np.array_equal(plan.grad_srcv(), s0)
    plan.close()


@pytest.mark.parametrize("variant", [1, 0])
def test_elastic_is_deterministic_at_full_size(A, ctx, variant):
    """el_sigma_fwd / el_vel_fwd (fields) and el_vel_adj / el_sigma_adj (source-only and material gradients) on the C5
    grid: repeat runs are bit-identical (same TMA ring protocol as the acoustic kernels, 2-3 CTAs/SM)."""
    NX = NY = 2000
    NSTEP = 40
    rng = np.random.default_rng(5)
    p = A.ElasticPropagatorParams(NX=NX, NY=NY, NSTEP=NSTEP, DELTAX=10.0, DELTAY=10.0, DELTAT=1e-3, vp_ref=3000.0,
                                  f0=10.0, variant=variant)
    shape = p.model_shape()
    vp = _layered(shape, 2500.0, 3500.0, rng)
    vs, rho = vp / 1.732, np.full(shape, 2500.0) * (1 + 0.02 * rng.random(shape))
    lam, mu = rho * (vp * vp - 2 * vs * vs), rho * vs * vs
    srci, srcj, srctype = np.array([1000, 700, 1200]), np.array([1000, 900, 1100]), np.array([2, 0, 4])
    srcv = np.stack([A.Ricker(p, 20.0 + k, 30.0, 1e6) for k in range(3)], 1)[:NSTEP]
    nrcv = 96
    rcvi, rcvj, rcvtype = np.linspace(900, 1100, nrcv).astype(np.int64), np.full(nrcv, 1010), np.arange(nrcv) % 5
    plan = A.ElasticPlan(p, srci, srcj, srctype, rcvi, rcvj, rcvtype, ctx=ctx)
    plan.set_model(rho, lam, mu); plan.set_srcv(srcv)
    plan.forward()
    r0, f0 = plan.rcvv(), [plan.snapshot(f, NSTEP) for f in (0, 4)]
    plan.set_obs(0.5 * r0)
    plan.gradient(True)
    ref = (plan.loss(), plan.grad_srcv(), plan.grad_rho(), plan.grad_lambda(), plan.grad_mu())
    assert np.abs(ref[4]).max() > 0
    for _ in range(4):
        plan.forward()
        assert np.array_equal(plan.rcvv(), r0)
        assert all(np.array_equal(plan.snapshot(f, NSTEP), x) for f, x in zip((0, 4), f0))
        plan.gradient(True)
        got = (plan.loss(), plan.grad_srcv(), plan.grad_rho(), plan.grad_lambda(), plan.grad_mu())
        assert got[0] == ref[0] and all(np.array_equal(a, b) for a, b in zip(got[1:], ref[1:]))
    plan.gradient(False)
    s1 = plan.grad_srcv()
    for _ in range(3):
        plan.gradient(False)
        assert np.array_equal(plan.grad_srcv(), s1)
    plan.close()
