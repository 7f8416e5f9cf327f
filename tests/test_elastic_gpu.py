"""GPU parity tests for the elastic path (both reference variants), through the C ABI, against the CPU oracle and
the golden vectors (torch-autograd restatement of the reference graph).  Forward traces / fields are bit-identical
(same expression order, no FMA); gradients <= 1e-10 relative (BASELINE.json), typically 1e-13."""
import numpy as np
import pytest

from conftest import golden, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _unpad(variant, a):
    return a if variant == 0 else a[..., 2:-2, 2:-2]


def _params(A, G):
    return A.ElasticPropagatorParams(NX=int(G["NX"]), NY=int(G["NY"]), NSTEP=int(G["NSTEP"]), DELTAX=float(G["dx"]),
                                     DELTAY=float(G["dy"]), DELTAT=float(G["dt"]), NPOINTS_PML=int(G["npml"]),
                                     vp_ref=float(G["vp_ref"]), ALPHA_MAX_PML=float(G["alpha_max"]),
                                     variant=int(G["variant"]))


@pytest.mark.parametrize("name", ["elastic_S.npz", "elastic_M.npz"])
def test_golden(A, ctx, name):
    G = golden(name)
    v = int(G["variant"])
    p = _params(A, G)
    src = A.ElasticSource(G["srci"], G["srcj"], G["srctype"], G["srcv"])
    rcv = A.ElasticReceiver(G["rcvi"], G["rcvj"], G["rcvtype"])
    rho, lam, mu = (_unpad(v, G[k]) for k in ("rho", "lam", "mu"))
    R = A.elastic_misfit_grad(p, src, rho, lam, mu, rcv, G["obs"], ctx=ctx)
    assert relerr(R["rcvv"], G["rcvv"]) < 1e-14
    assert abs(R["loss"] - float(G["loss"])) / float(G["loss"]) < 1e-12
    assert relerr(R["grad_srcv"], G["grad_srcv"]) < TOL
    for k, gk in (("grad_rho", "grad_rho"), ("grad_lambda", "grad_lam"), ("grad_mu", "grad_mu")):
        assert relerr(R[k], _unpad(v, G[gk])) < TOL, k
    # source-time-function-only gradient (no forward history at all) gives the same grad_srcv
    R2 = A.elastic_misfit_grad(p, src, rho, lam, mu, rcv, G["obs"], material_grads=False, ctx=ctx)
    assert relerr(R2["grad_srcv"], G["grad_srcv"]) < TOL and R2["loss"] == R["loss"]


def _case(po, rng, variant, NX, NY, NSTEP, npml=10):
    h, dt = 1.0, 1e-4
    H, W = po.elastic_dims(variant, NX, NY)
    ax, bx = po.elastic_cpml_1d(NX, h, dt, npml=npml, vp_ref=3300.0, alpha_max=np.pi * 15)
    ay, by = po.elastic_cpml_1d(NY, h, dt, npml=npml, vp_ref=3300.0, alpha_max=np.pi * 15)
    vp = 3000.0 * (1 + 0.1 * rng.random((H, W)))
    vs = vp / 1.732 * (1 + 0.05 * rng.random((H, W)))
    rho = 2800.0 * (1 + 0.1 * rng.random((H, W)))
    mu, lam = rho * vs * vs, rho * (vp * vp - 2 * vs * vs)
    nsrc = 7
    srci = rng.integers(3, NX - 2, nsrc); srcj = rng.integers(3, NY - 2, nsrc)
    srctype = np.array([0, 1, 2, 3, 4, 2, 0])
    srci[5], srcj[5] = srci[2], srcj[2]       # two sxx sources on one cell
    srcv = np.stack([po.ricker(NSTEP, 6.0 + k, 12.0 + k, 1e3 * (1 + k)) for k in range(nsrc)], 1)
    nrcv = 30
    rcvi = rng.integers(1, NX + 1, nrcv); rcvj = rng.integers(1, NY + 1, nrcv)
    rcvtype = rng.integers(0, 5, nrcv)
    rcvi[0], rcvj[0], rcvtype[0] = srci[2], srcj[2], 2   # stress receiver on a stress-source cell
    rcvi[1], rcvj[1], rcvtype[1] = srci[0], srcj[0], 0   # velocity receiver on a velocity-source cell
    rcvi[2], rcvj[2], rcvtype[2] = rcvi[0], rcvj[0], 2   # duplicate receiver
    return dict(h=h, dt=dt, ax=ax, bx=bx, ay=ay, by=by, rho=rho, lam=lam, mu=mu, srci=srci, srcj=srcj,
                srctype=srctype, srcv=srcv, rcvi=rcvi, rcvj=rcvj, rcvtype=rcvtype, npml=npml)


@pytest.fixture(params=["march", "generic"])
def tiling(request, monkeypatch):
    """Both CTA kinds of the step kernels: `march` forces the TMA-ring marching CTAs onto the (small) test grids'
    PML-free box, `generic` keeps every cell on the one-cell-per-thread path (ADSEIS_EL_MARCH_MIN = minimum box
    cells for marching; by default small grids stay generic)."""
    monkeypatch.setenv("ADSEIS_EL_MARCH_MIN", "0" if request.param == "march" else str(1 << 40))
    return request.param


@pytest.mark.parametrize("variant,shape", [(0, (150, 170, 40)), (1, (130, 200, 36)), (0, (500, 500, 12)),
                                           (1, (90, 1200, 10))])
def test_vs_oracle(A, ctx, po, variant, shape, tiling):
    NX, NY, NSTEP = shape
    rng = np.random.default_rng(100 * variant + NX)
    K = _case(po, rng, variant, NX, NY, NSTEP)
    args = (variant, NX, NY, NSTEP, K["dt"], K["h"], K["h"], K["ax"], K["bx"], K["ay"], K["by"], K["rho"], K["lam"],
            K["mu"], K["srci"], K["srcj"], K["srctype"], K["srcv"], K["rcvi"], K["rcvj"], K["rcvtype"])
    r0, hist0 = po.elastic_forward(*args, want_hist=True)
    obs = 0.6 * r0 + 0.05 * np.abs(r0).max() * rng.standard_normal(r0.shape)
    O = po.elastic_misfit_grad(*args, obs)
    p = A.ElasticPropagatorParams(NX=NX, NY=NY, NSTEP=NSTEP, DELTAX=K["h"], DELTAY=K["h"], DELTAT=K["dt"],
                                  NPOINTS_PML=K["npml"], vp_ref=3300.0, ALPHA_MAX_PML=np.pi * 15, variant=variant)
    rho, lam, mu = (_unpad(variant, K[k]) for k in ("rho", "lam", "mu"))
    plan = A.ElasticPlan(p, K["srci"], K["srcj"], K["srctype"], K["rcvi"], K["rcvj"], K["rcvtype"], ctx=ctx)
    plan.set_model(rho, lam, mu); plan.set_srcv(K["srcv"])
    plan.forward()
    assert np.array_equal(plan.rcvv(), r0)
    for f in range(5):   # post-injection fields of resident slots, bit-identical
        for s in (1, NSTEP // 2, NSTEP):
            assert np.array_equal(plan.snapshot(f, s), _unpad(variant, hist0[f, s])), (f, s)
    plan.set_obs(obs)
    plan.gradient(True)
    assert np.array_equal(plan.rcvv(), r0)
    assert abs(plan.loss() - O["loss"]) / O["loss"] < 1e-12
    assert relerr(plan.grad_srcv(), O["grad_srcv"]) < TOL
    assert relerr(plan.grad_rho(), _unpad(variant, O["grad_rho"])) < TOL
    assert relerr(plan.grad_lambda(), _unpad(variant, O["grad_lam"])) < TOL
    assert relerr(plan.grad_mu(), _unpad(variant, O["grad_mu"])) < TOL
    plan.gradient(False)
    assert relerr(plan.grad_srcv(), O["grad_srcv"]) < TOL
    plan.close()


def test_checkpointed_gradient_equals_full_history(A, ctx, po, tiling):
    rng = np.random.default_rng(9)
    variant, NX, NY, NSTEP = 0, 120, 140, 30
    K = _case(po, rng, variant, NX, NY, NSTEP)
    p = A.ElasticPropagatorParams(NX=NX, NY=NY, NSTEP=NSTEP, DELTAX=K["h"], DELTAY=K["h"], DELTAT=K["dt"],
                                  NPOINTS_PML=K["npml"], vp_ref=3300.0, ALPHA_MAX_PML=np.pi * 15, variant=variant)
    obs = np.zeros((len(K["rcvi"]), NSTEP + 1))
    out = []
    slot_bytes = None
    for budget in (0, "small"):
        plan = A.ElasticPlan(p, K["srci"], K["srcj"], K["srctype"], K["rcvi"], K["rcvj"], K["rcvtype"], ctx=ctx,
                             hist_bytes_budget=0 if budget == 0 else 7 * slot_bytes)
        slot_bytes = plan.info()["slot_doubles"] * 8
        plan.set_model(K["rho"], K["lam"], K["mu"]); plan.set_srcv(K["srcv"]); plan.set_obs(obs)
        plan.gradient(True)
        out.append((plan.loss(), plan.grad_rho(), plan.grad_lambda(), plan.grad_mu(), plan.grad_srcv(), plan.info()))
        plan.close()
    assert out[0][5]["segments"] == 1 and out[1][5]["segments"] >= 4 and out[1][5]["recomputed_steps"] > 0
    assert out[0][0] == out[1][0]
    for k in range(1, 5):
        assert np.array_equal(out[0][k], out[1][k])


def test_api_names(A, ctx, po):
    """ElasticPropagatorSolver / SimulatedObservation! as in examples/demo/ElasticWave.jl:8-27 (scaled down)."""
    p = A.ElasticPropagatorParams(NX=60, NY=60, NSTEP=40, DELTAT=1e-4, DELTAX=1.0, DELTAY=1.0, vp_ref=3300.0)
    source = A.Ricker(p, 15.0, 100.0, 1e6)
    src = A.ElasticSource([p.NX // 2], [p.NY // 2], [0], source.reshape(-1, 1))
    lam, mu, rho = A.compute_lame_parameters(p.NX, p.NY, 3000.0, 3000.0 / 1.732, 2800.0)
    model = A.ElasticPropagatorSolver(p, src, rho, lam, mu, ctx=ctx)
    rcv = A.ElasticReceiver([20, 25, 30], [30, 30, 30], [0, 1, 2])
    A.SimulatedObservation_(model, rcv)
    assert rcv.rcvv.shape == (3, 41)
    vx = model.vx
    assert vx.shape == (41, 62, 62)
    assert np.array_equal(rcv.rcvv[0], vx[:, 19, 29])
    ax, bx, kx, ay, by, ky = A.compute_PML_Params(p)
    a0, b0 = po.elastic_cpml_1d(60, 1.0, 1e-4, npml=12, vp_ref=3300.0, alpha_max=2 * np.pi * 2.5)
    assert np.array_equal(ax.reshape(-1), a0) and np.array_equal(bx.reshape(-1), b0)
