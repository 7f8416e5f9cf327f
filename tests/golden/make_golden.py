"""Generate the golden vectors under tests/golden/ (run in the BUILD container, where /root/reference exists).

The reference ships no golden vectors and cannot run here (Julia + ADCME + TensorFlow 1.15 are absent), so the
vectors are produced by the strongest stand-ins available:
  acoustic_*.npz : the reference's OWN C++ op bodies (AcousticOneStepCpu.h, ScatterAddOps.h) driven in the order of
                   src/Core.jl:562-620 by oracle/ref_shim.cpp (oracle/_ref/libadseis_ref.so)
  elastic_*.npz, acoustic_kernel0.npz : the reference's elastic / PropagatorKernel=0 TensorFlow graphs
                   (src/Core.jl:31-228, 528-620; src/MPIElastic.jl:374-645) recorded op by op and executed /
                   differentiated with the reference's OWN gather, scatter_add, scatter_nd, add_source and get_receive
                   bodies (oracle/ref_graph.inc -> oracle/_ref/libadseis_ref.so); only the element-wise fp64
                   arithmetic in between is restated.  (Round 1 generated them from torch restatements,
                   oracle/torch_*.py, which are kept as an independent cross-check.)
Usage: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def acoustic_case(name, NX, NY, NSTEP, dx, dy, dt, npml, vp_ref, seed, use=(True, True, True, True)):
    rng = np.random.default_rng(seed)
    sig, tau = po.acoustic_pml(NX, NY, dx, dy, npml=npml, vp_ref=vp_ref, use=use)
    c = vp_ref * (1 + 0.1 * rng.standard_normal((NX + 2, NY + 2)))
    srci = np.array([NX // 2, 3, NX // 2, 1], dtype=np.int64)        # duplicate cell + a source on the ring row
    srcj = np.array([NY // 2, 4, NY // 2, NY // 3], dtype=np.int64)
    srcv = np.stack([po.ricker(NSTEP, 10., 30., 1e6), po.ricker(NSTEP, 8., 40., 5e5), po.ricker(NSTEP, 12., 20., 2e5),
                     po.ricker(NSTEP, 9., 25., 3e5)], 1)
    rcvi = np.concatenate([np.arange(2, NX, 2), [NX // 2, NX // 2]]).astype(np.int64)  # incl. duplicates
    rcvj = np.concatenate([np.full(len(np.arange(2, NX, 2)), 4), [NY // 2, NY // 2]]).astype(np.int64)
    u, rcvv = po.acoustic_forward(NX, NY, NSTEP, dt, dx, dy, sig, tau, c, srci, srcj, srcv, rcvi, rcvj, which="ref")
    obs = rcvv * (1 + 0.1 * rng.standard_normal(rcvv.shape)) + 0.01 * np.abs(rcvv).max() * rng.standard_normal(rcvv.shape)
    loss, gc, gs = po.acoustic_misfit_grad(NX, NY, NSTEP, dt, dx, dy, sig, tau, c, srci, srcj, rcvi, rcvj, obs, u,
                                           which="ref")
    np.savez_compressed(os.path.join(HERE, name), NX=NX, NY=NY, NSTEP=NSTEP, dx=dx, dy=dy, dt=dt, npml=npml,
                        vp_ref=vp_ref, use=np.array(use), c=c, srci=srci, srcj=srcj, srcv=srcv, rcvi=rcvi, rcvj=rcvj,
                        obs=obs, rcvv=rcvv, u_last=u[-1], u_mid=u[NSTEP // 2], loss=loss, grad_c=gc, grad_srcv=gs,
                        sigx=sig.reshape(NX + 2, NY + 2)[:, 0], tauy=tau.reshape(NX + 2, NY + 2)[0, :])
    print(name, "loss", loss, "|grad_c|max", np.abs(gc).max())


def acoustic_step_case(name, seed):
    """The inputs of deps/CustomOps/AcousticOneStepCpu/gradtest.jl:15-31 (nx=ny=10, h=dt=0.1, uniform(0,1))."""
    rng = np.random.default_rng(seed)
    NX = NY = 10
    N = (NX + 2) * (NY + 2)
    ins = [rng.random(N) for _ in range(7)]
    g = [rng.random(N) for _ in range(3)]
    fwd = po.acoustic_step_fwd(*ins, 0.1, 0.1, 0.1, NX, NY, "ref")
    bwd = po.acoustic_step_bwd(*g, ins[0], ins[4], ins[5], ins[6], 0.1, 0.1, 0.1, NX, NY, "ref")
    np.savez_compressed(os.path.join(HERE, name), ins=np.stack(ins), g=np.stack(g), fwd=np.stack(fwd),
                        bwd=np.stack(bwd))
    print(name, "ok")


def elastic_case(name, variant, NX, NY, NSTEP, seed):
    rng = np.random.default_rng(seed)
    dx = dy = 1.0
    dt = 1e-4
    H, W = po.elastic_dims(variant, NX, NY)
    kw = dict(npml=5, vp_ref=3300., alpha_max=np.pi * 15)
    ax, bx = po.elastic_cpml_1d(NX, dx, dt, **kw)
    ay, by = po.elastic_cpml_1d(NY, dy, dt, **kw)
    vp = 3000. * (1 + 0.1 * rng.random((H, W)))
    vs = vp / 1.732 * (1 + 0.05 * rng.random((H, W)))
    rho = 2800. * (1 + 0.1 * rng.random((H, W)))
    mu = rho * vs * vs
    lam = rho * (vp * vp - 2 * vs * vs)
    srci = np.array([NX // 2, 3, NX // 2, 7, 9, NX - 1, NX // 2])
    srcj = np.array([NY // 2, 4, NY // 2, 8, 3, NY - 2, NY // 2])
    srctype = np.array([0, 1, 2, 3, 4, 2, 0])
    srcv = rng.standard_normal((NSTEP, len(srci))) * 1e3
    rcvi = np.array([2, 5, 8, 11, 14, NX // 2, 3, NX // 2])
    rcvj = np.array([3, 3, 6, 9, 12, NY // 2, 4, NY // 2])
    rcvtype = np.array([0, 1, 2, 3, 4, 2, 1, 2])
    r0, _ = po.elastic_forward(variant, NX, NY, NSTEP, dt, dx, dy, ax, bx, ay, by, rho, lam, mu, srci, srcj, srctype,
                               srcv, rcvi, rcvj, rcvtype)
    obs = r0 * (1 + 0.2 * rng.standard_normal(r0.shape)) + 0.05 * np.abs(r0).max() * rng.standard_normal(r0.shape)
    B = po.ref_elastic(variant, NX, NY, NSTEP, dt, dx, dy, ax, bx, ay, by, rho, lam, mu, srci, srcj, srctype, srcv,
                       rcvi, rcvj, rcvtype, obs)
    np.savez_compressed(os.path.join(HERE, name), variant=variant, NX=NX, NY=NY, NSTEP=NSTEP, dx=dx, dy=dy, dt=dt,
                        npml=5, vp_ref=3300., alpha_max=np.pi * 15, ax=ax, bx=bx, ay=ay, by=by, rho=rho, lam=lam,
                        mu=mu, srci=srci, srcj=srcj, srctype=srctype, srcv=srcv, rcvi=rcvi, rcvj=rcvj, rcvtype=rcvtype,
                        obs=obs, rcvv=B["rcvv"], loss=B["loss"], grad_rho=B["grad_rho"], grad_lam=B["grad_lam"],
                        grad_mu=B["grad_mu"], grad_srcv=B["grad_srcv"])
    print(name, "loss", B["loss"])


def acoustic_kernel0_case(name, NX, NY, NSTEP, seed):
    """PropagatorKernel=0 (src/Core.jl:528-549): golden values from the graph over the reference's own gather /
    scatter_nd / scatter_add op bodies (oracle/ref_graph.inc)."""
    rng = np.random.default_rng(seed)
    dx, dy, dt, npml, vp_ref = 10.0, 8.0, 1e-3, 6, 2500.0
    sig, tau = po.acoustic_pml(NX, NY, dx, dy, npml=npml, vp_ref=vp_ref)
    c = vp_ref * (1 + 0.1 * rng.random((NX + 2, NY + 2)))
    # interior source, two sources inside the absorbing frame on neighbouring cells, one on the ring, one in a corner
    srci = np.array([NX // 2, 3, 4, 1, NX - 1], dtype=np.int64)
    srcj = np.array([NY // 2, 5, 5, NY // 3, NY], dtype=np.int64)
    srcv = np.stack([po.ricker(NSTEP, 6.0 + k, 10.0 + k, 1e6) for k in range(5)], 1)
    rcvi = rng.integers(1, NX + 3, 24)
    rcvj = rng.integers(1, NY + 3, 24)
    r0 = po.ref_acoustic_graph(0, NX, NY, NSTEP, dt, dx, dy, sig, tau, c, srci, srcj, srcv, rcvi, rcvj)["rcvv"]
    obs = 0.7 * r0 + 0.01 * np.abs(r0).max() * rng.standard_normal(r0.shape)
    R = po.ref_acoustic_graph(0, NX, NY, NSTEP, dt, dx, dy, sig, tau, c, srci, srcj, srcv, rcvi, rcvj, obs)
    np.savez_compressed(os.path.join(HERE, name), NX=NX, NY=NY, NSTEP=NSTEP, dx=dx, dy=dy, dt=dt, npml=npml,
                        vp_ref=vp_ref, c=c, srci=srci, srcj=srcj, srcv=srcv, rcvi=rcvi, rcvj=rcvj, obs=obs,
                        rcvv=R["rcvv"], loss=R["loss"], grad_c=R["grad_c"], grad_srcv=R["grad_srcv"])
    print(name, "loss", R["loss"], "|grad_c|max", np.abs(R["grad_c"]).max())


def marmousi_case(name, nstep=400, shot=3):
    """The reference's own model fixture (examples/nn_fwi/models/marmousi2-model-true.mat: 202 x 68 padded cells, 8
    shots, 183 receivers; the FWI scripts run it with AcousticPropagatorSolver, examples/nn_fwi/FWI_inversion.jl): one
    shot, the first `nstep` of its 1678 time steps, through the reference's C++ op bodies; observed data from the
    smooth starting model of the same directory, exactly the misfit the inversion scripts start from."""
    import adseis_b200 as A
    d = "/root/reference/examples/nn_fwi/models/"
    param, vp_true = A.io.load_acoustic_model(d + "marmousi2-model-true.mat")
    _, vp_smooth = A.io.load_acoustic_model(d + "marmousi2-model-smooth.mat")
    src = A.io.load_acoustic_source(d + "marmousi2-model-true.mat")[shot]
    rcv = A.io.load_acoustic_receiver(d + "marmousi2-model-true.mat")[shot]
    NX, NY, dx, dy, dt = param.NX, param.NY, param.DELTAX, param.DELTAY, param.DELTAT
    vp_ref = float(vp_true.mean())
    srcv = np.ascontiguousarray(src.srcv[:nstep])
    sig, tau = po.acoustic_pml(NX, NY, dx, dy, npml=param.NPOINTS_PML, vp_ref=vp_ref)
    _, obs = po.acoustic_forward(NX, NY, nstep, dt, dx, dy, sig, tau, vp_true, src.srci, src.srcj, srcv, rcv.rcvi,
                                 rcv.rcvj, which="ref")
    u, rcvv = po.acoustic_forward(NX, NY, nstep, dt, dx, dy, sig, tau, vp_smooth, src.srci, src.srcj, srcv, rcv.rcvi,
                                  rcv.rcvj, which="ref")
    loss, gc, gs = po.acoustic_misfit_grad(NX, NY, nstep, dt, dx, dy, sig, tau, vp_smooth, src.srci, src.srcj,
                                           rcv.rcvi, rcv.rcvj, obs, u, which="ref")
    np.savez_compressed(os.path.join(HERE, name), NX=NX, NY=NY, NSTEP=nstep, dx=dx, dy=dy, dt=dt,
                        npml=param.NPOINTS_PML, vp_ref=vp_ref, c=vp_smooth, srci=src.srci, srcj=src.srcj, srcv=srcv,
                        rcvi=rcv.rcvi, rcvj=rcv.rcvj, obs=obs, rcvv=rcvv, loss=loss, grad_c=gc, grad_srcv=gs)
    print(name, "loss", loss, "|grad_c|max", np.abs(gc).max(), "|rcvv|max", np.abs(rcvv).max())


if __name__ == "__main__":
    assert po.has_ref(), "oracle/_ref is not built: run `make -C oracle` where /root/reference exists"
    acoustic_step_case("acoustic_step_gradtest.npz", 233)
    acoustic_case("acoustic_small.npz", 40, 30, 120, 10.0, 10.0, 1e-3, 6, 2000.0, 1234)
    acoustic_case("acoustic_nopml_y.npz", 33, 45, 60, 8.0, 12.0, 1e-3, 5, 2500.0, 99, use=(True, True, False, False))
    elastic_case("elastic_S.npz", 0, 26, 22, 30, 7)
    elastic_case("elastic_M.npz", 1, 26, 22, 30, 8)
    marmousi_case("acoustic_marmousi2_shot3.npz")
    acoustic_kernel0_case("acoustic_kernel0.npz", 30, 37, 60, 17)
