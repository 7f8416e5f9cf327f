"""CPU tests of the data formats either side of the hot path: the MAT model/shot schema of
src/IO.jl:4-161 round-trips through adseis_b200.io, the parameterisation helpers of fwi.py reproduce
src/IO.jl:172-197 / src/Utils.jl:603-632,237-243, and the oracle reproduces the golden vector generated from the
reference's own Marmousi fixture bit for bit."""
import numpy as np
import pytest

from conftest import golden


def test_mat_roundtrip_acoustic(A, tmp_path):
    rng = np.random.default_rng(3)
    nx, ny, nt = 40, 30, 50
    vp = 2000 + 500 * rng.random((nx, ny))
    srcs = [A.AcousticSource([5 + k], [7], rng.standard_normal((nt, 1))) for k in range(3)]
    srcs.append(A.AcousticSource([3, 4], [9, 10], rng.standard_normal((nt, 2))))     # a two-point shot
    rcvs = [A.AcousticReceiver(np.arange(2, 30), np.full(28, 4 + k)) for k in range(4)]
    fn = str(tmp_path / "m.mat")
    A.io.save_model(fn, vp, 0.5 * vp, 0 * vp + 2500, srcs, rcvs, 10.0, 12.0, 1e-3, nt, 5.0)
    p, vp2 = A.io.load_acoustic_model(fn, IT_DISPLAY=0)
    assert (p.NX, p.NY, p.NSTEP) == (nx - 2, ny - 2, nt) and (p.DELTAX, p.DELTAY, p.DELTAT) == (10.0, 12.0, 1e-3)
    assert np.array_equal(vp2, vp)
    s2, r2 = A.io.load_acoustic_source(fn), A.io.load_acoustic_receiver(fn)
    assert len(s2) == 4 and len(r2) == 4
    for a, b in zip(srcs, s2):
        assert np.array_equal(a.srci, b.srci) and np.array_equal(a.srcj, b.srcj) and np.array_equal(a.srcv, b.srcv)
    for a, b in zip(rcvs, r2):
        assert np.array_equal(a.rcvi, b.rcvi) and np.array_equal(a.rcvj, b.rcvj)


def test_mat_roundtrip_elastic(A, tmp_path):
    rng = np.random.default_rng(4)
    nx, ny, nt = 36, 28, 20
    vp = 3000 + 100 * rng.random((nx, ny))
    srcs = [A.ElasticSource([6, 6], [8, 8], [2, 3], rng.standard_normal((nt, 2)))]
    rcvs = [A.ElasticReceiver(np.arange(3, 20), np.full(17, 5), np.zeros(17, dtype=np.int64))]
    fn = str(tmp_path / "e.mat")
    A.io.save_model(fn, vp, vp / 1.7, 0 * vp + 2200, srcs, rcvs, 8.0, 8.0, 5e-4, nt, 10.0)
    p, vp2, vs2, rho2 = A.io.load_elastic_model(fn)
    assert (p.NX, p.NY, p.NSTEP) == (nx - 2, ny - 2, nt)
    assert p.vp_ref == pytest.approx(vp.mean()) and p.f0 == 5.0          # src/IO.jl:30-48: f0 / 2, vp_ref = mean
    assert np.array_equal(vs2, vp / 1.7) and np.array_equal(rho2, 0 * vp + 2200)
    s2, r2 = A.io.load_elastic_source(fn)[0], A.io.load_elastic_receiver(fn)[0]
    assert np.array_equal(s2.srctype, [2, 3]) and np.array_equal(s2.srcv, srcs[0].srcv)
    assert np.array_equal(r2.rcvi, rcvs[0].rcvi) and np.array_equal(r2.rcvtype, rcvs[0].rcvtype)
    pm = A.io.load_params(fn, "MPIElastic")
    assert pm.variant == 1 and (pm.NX, pm.NY) == (nx, ny)


def test_oracle_reproduces_marmousi_golden(po):
    """The golden file was made by the reference's C++ op bodies on the reference's own Marmousi fixture (shot 3 of
    examples/nn_fwi/models/marmousi2-model-true.mat, 400 steps, smooth starting model): the C oracle matches it bit
    for bit."""
    G = golden("acoustic_marmousi2_shot3.npz")
    NX, NY, NSTEP = int(G["NX"]), int(G["NY"]), int(G["NSTEP"])
    dx, dy, dt = float(G["dx"]), float(G["dy"]), float(G["dt"])
    sig, tau = po.acoustic_pml(NX, NY, dx, dy, npml=int(G["npml"]), vp_ref=float(G["vp_ref"]))
    u, rcvv = po.acoustic_forward(NX, NY, NSTEP, dt, dx, dy, sig, tau, G["c"], G["srci"], G["srcj"], G["srcv"],
                                  G["rcvi"], G["rcvj"])
    assert np.array_equal(rcvv, G["rcvv"])
    loss, gc, gs = po.acoustic_misfit_grad(NX, NY, NSTEP, dt, dx, dy, sig, tau, G["c"], G["srci"], G["srcj"],
                                           G["rcvi"], G["rcvj"], G["obs"], u)
    assert loss == float(G["loss"]) and np.array_equal(gc, G["grad_c"]) and np.array_equal(gs, G["grad_srcv"])


def test_parameterisation_helpers(A):
    import torch
    fwi = A.fwi
    x = np.linspace(1500, 3500, 12).reshape(3, 4)
    cv = fwi.ConstantOrVariable(x, trainable=False)
    assert np.array_equal(cv().numpy(), x) and not list(cv.parameters())
    mask = np.zeros((3, 4)); mask[1:, :] = 1
    tv = fwi.ConstantOrVariable(x, trainable=True, mask=mask)
    assert np.allclose(tv().detach().numpy(), x, rtol=1e-15)
    with torch.no_grad():
        tv.x_ += 0.1
    y = tv().detach().numpy()
    assert np.allclose(y[0], x[0], rtol=1e-15) and np.allclose(y[1:], x[1:] + 0.1 * x.mean(), rtol=1e-14)
    lam, mu, rho = fwi.compute_properties(*(torch.tensor(v, dtype=torch.float64) for v in (3000.0, 1500.0, 2000.0)))
    assert float(lam) == 2000.0 * (3000.0 ** 2 - 2 * 1500.0 ** 2) and float(mu) == 2000.0 * 1500.0 ** 2
    p = A.AcousticPropagatorParams(NX=6, NY=5, NSTEP=4, DELTAX=10.0, DELTAY=10.0)
    xs = torch.tensor(30.0, dtype=torch.float64, requires_grad=True)
    srci, srcj, srcv = fwi.variable_source(p, xs, 20.0, np.array([1.0, 2.0, 3.0, 4.0]))
    assert srcv.shape == (4, 8 * 7) and len(srci) == 56
    k = int(np.argmax(srcv[0].detach().numpy()))
    assert (srci[k], srcj[k]) == (4, 3)                 # the blob peaks at x = (srci-1) dx = 30, y = 20
    assert float(srcv[0, k].detach()) == pytest.approx(1 / (2 * np.pi))
    srcv.sum().backward()
    assert xs.grad is not None


class _QuadraticPlan:
    """Stand-in for AcousticPlan / ElasticPlan in CPU tests of the autograd wiring: loss = sum w (c - c*)^2 (+ a term in
    srcv), gradients analytic.  (The real plans need a GPU; tests/test_fwi_gpu.py covers them.)"""

    def __init__(self, target, weight, nstep=5, nsrc=2):
        self.t, self.w = np.asarray(target, dtype=np.float64), np.asarray(weight, dtype=np.float64)
        self.model_shape, self.nstep, self.nsrc = self.t.shape, nstep, nsrc
        self.c, self.s, self.calls = None, np.zeros((nstep, nsrc)), 0

        class _Ctx:
            def sync(self_inner):
                pass
        self.ctx = _Ctx()

    def set_model(self, *arrs):
        self.c = sum(np.array(a, dtype=np.float64).reshape(self.model_shape) for a in arrs)

    def set_srcv(self, s, rows=None):
        self.s = np.array(s, dtype=np.float64)[:self.nstep]

    def gradient(self, material_grads=True):
        self.calls += 1

    def loss(self):
        return float((self.w * (self.c - self.t) ** 2).sum() + 0.5 * (self.s ** 2).sum())

    def grad_c(self, out=None):
        g = 2 * self.w * (self.c - self.t)
        if out is not None:
            out[...] = g
        return g

    grad_rho = grad_lambda = grad_mu = lambda self: 2 * self.w * (self.c - self.t)

    def grad_srcv(self):
        return self.s.copy()


def test_autograd_wiring_and_lbfgs_on_cpu(A):
    import torch
    fwi = A.fwi
    rng = np.random.default_rng(8)
    target, weight = 2000 + 500 * rng.random((6, 5)), 0.5 + rng.random((6, 5))
    plan = _QuadraticPlan(target, weight)
    mask = np.ones((6, 5)); mask[:, 0] = 0
    x0 = np.full((6, 5), 2100.0)
    vp = fwi.ConstantOrVariable(x0, trainable=True, mask=mask)
    srcv = torch.tensor(rng.standard_normal((7, 2)), requires_grad=True)      # 7 rows, the plan uses 5
    loss = fwi.acoustic_misfit(plan, vp(), srcv)
    (2.0 * loss).backward()
    want = 2.0 * 2 * weight * (x0 - target) * mask * x0.mean()
    assert np.allclose(vp.x_.grad.numpy(), want, rtol=1e-13)
    g = srcv.grad.numpy()
    assert np.allclose(g[:5], 2.0 * srcv.detach().numpy()[:5], rtol=1e-14) and not g[5:].any()
    # L-BFGS drives the masked quadratic to its minimum; unmasked column stays at the start value
    vp2 = fwi.ConstantOrVariable(x0, trainable=True, mask=mask)
    plan = _QuadraticPlan(target, weight)        # fresh: no source term
    losses = fwi.LBFGS_(lambda: fwi.acoustic_misfit(plan, vp2()), vp2.parameters(), max_iter=40)
    out = vp2().detach().numpy()
    assert losses[-1] < 1e-12 * losses[0] + float((weight[:, 0] * (x0[:, 0] - target[:, 0]) ** 2).sum()) * (1 + 1e-9)
    assert np.allclose(out[:, 1:], target[:, 1:], rtol=1e-5) and np.allclose(out[:, 0], x0[:, 0], rtol=1e-15)
    assert all(b <= a * (1 + 1e-12) for a, b in zip(losses, losses[1:]))
    # elastic wrapper: three material tensors, gradients routed to each
    rho, lam, mu = (torch.tensor(rng.random((6, 5)) * 700, requires_grad=True) for _ in range(3))
    fwi.elastic_misfit(plan, rho, lam, mu).backward()
    gsum = 2 * weight * ((rho + lam + mu).detach().numpy() - target)
    for t in (rho, lam, mu):
        assert np.allclose(t.grad.numpy(), gsum, rtol=1e-13)
