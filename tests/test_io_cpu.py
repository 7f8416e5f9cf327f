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
    assert float(srcv[0, k]) == pytest.approx(1 / (2 * np.pi))
    srcv.sum().backward()
    assert xs.grad is not None
