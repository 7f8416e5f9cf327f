"""ctypes/numpy front-end for the CPU checkers -- TEST INFRASTRUCTURE ONLY.

`oracle`  = oracle/liboracle.so      (plain-C restatement, oracle/oracle.c)
`ref`     = oracle/_ref/libadseis_ref.so (the reference's own C++ op bodies, oracle/ref_shim.cpp), may be absent.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product package
(adseismic.jl_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_i64 = C.c_longlong
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(_i64)


def build(quiet=True):
    """(Re)build liboracle.so and, when /root/reference is present, _ref/libadseis_ref.so."""
    subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL if quiet else None)


def _load(path):
    return C.CDLL(path) if os.path.exists(path) else None


def lib():
    p = os.path.join(_HERE, "liboracle.so")
    if not os.path.exists(p):
        build()
    return C.CDLL(p)


def ref_lib():
    return _load(os.path.join(_HERE, "_ref", "libadseis_ref.so"))


def has_ref():
    return os.path.exists(os.path.join(_HERE, "_ref", "libadseis_ref.so"))


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(_ip)


# ----------------------------------------------------------------------------------------------
# input builders
# ----------------------------------------------------------------------------------------------
def ricker(nt, a, shift, amp=1.0):
    out = np.empty(nt)
    lib().orc_ricker(_i64(nt), C.c_double(a), C.c_double(shift), C.c_double(amp), out.ctypes.data_as(_dp))
    return out


def gauss(nt, dt, a, shift=None, amp=1.0):
    if shift is None:
        shift = 1.2 / a
    out = np.empty(nt)
    lib().orc_gauss(_i64(nt), C.c_double(dt), C.c_double(a), C.c_double(shift), C.c_double(amp),
                    out.ctypes.data_as(_dp))
    return out


def acoustic_pml_1d(NX, NY, dx, dy, npml=12, Rcoef=1e-3, vp_ref=1000.0, use=(True, True, True, True)):
    sx = np.empty(NX + 2)
    ty = np.empty(NY + 2)
    lib().orc_acoustic_pml_1d(_i64(NX), _i64(NY), C.c_double(dx), C.c_double(dy), _i64(npml), C.c_double(Rcoef),
                              C.c_double(vp_ref), *[C.c_int(int(u)) for u in use], sx.ctypes.data_as(_dp),
                              ty.ctypes.data_as(_dp))
    return sx, ty


def acoustic_pml(NX, NY, dx, dy, npml=12, Rcoef=1e-3, vp_ref=1000.0, use=(True, True, True, True)):
    N = (NX + 2) * (NY + 2)
    s = np.empty(N)
    t = np.empty(N)
    lib().orc_acoustic_pml(_i64(NX), _i64(NY), C.c_double(dx), C.c_double(dy), _i64(npml), C.c_double(Rcoef),
                           C.c_double(vp_ref), *[C.c_int(int(u)) for u in use], s.ctypes.data_as(_dp),
                           t.ctypes.data_as(_dp))
    return s, t


def elastic_cpml_1d(n, h, dt, npml=12, npower=2.0, kmax=1.0, alpha_max=2 * np.pi * 2.5, Rcoef=1e-3,
                    vp_ref=2000.0, use_min=True, use_max=True):
    a = np.empty(2 * n)
    b = np.empty(2 * n)
    lib().orc_elastic_cpml_1d(_i64(n), C.c_double(h), C.c_double(dt), _i64(npml), C.c_double(npower),
                              C.c_double(kmax), C.c_double(alpha_max), C.c_double(Rcoef), C.c_double(vp_ref),
                              C.c_int(int(use_min)), C.c_int(int(use_max)), a.ctypes.data_as(_dp),
                              b.ctypes.data_as(_dp))
    return a, b


# ----------------------------------------------------------------------------------------------
# acoustic
# ----------------------------------------------------------------------------------------------
def acoustic_step_fwd(w, wold, phi, psi, sigma, tau, c2, dt, hx, hy, NX, NY, which="oracle"):
    N = (NX + 2) * (NY + 2)
    u, phio, psio = np.empty(N), np.empty(N), np.empty(N)
    args = [_d(x)[1] for x in (w, wold, phi, psi, sigma, tau, c2)] + [C.c_double(dt), C.c_double(hx), C.c_double(hy),
                                                                   _i64(NX), _i64(NY)] + \
           [x.ctypes.data_as(_dp) for x in (u, phio, psio)]
    if which == "oracle":
        lib().orc_acoustic_step_fwd(*args)
    else:
        ref_lib().ref_op_acoustic_step_fwd(*args)
    return u, phio, psio


def acoustic_step_bwd(gu, gphio, gpsio, w, sigma, tau, c2, dt, hx, hy, NX, NY, which="oracle"):
    N = (NX + 2) * (NY + 2)
    outs = [np.empty(N) for _ in range(5)]
    op = [x.ctypes.data_as(_dp) for x in outs]
    sc = [C.c_double(dt), C.c_double(hx), C.c_double(hy), _i64(NX), _i64(NY)]
    if which == "oracle":
        lib().orc_acoustic_step_bwd(*op, _d(gu)[1], _d(gphio)[1], _d(gpsio)[1], _d(w)[1], _d(sigma)[1],
                                    _d(tau)[1], _d(c2)[1], *sc)
    else:
        z = np.zeros(N)
        ref_lib().ref_op_acoustic_step_bwd(*op, _d(gu)[1], _d(gphio)[1], _d(gpsio)[1], _d(w)[1], _d(z)[1],
                                           _d(z)[1], _d(z)[1], _d(sigma)[1], _d(tau)[1], _d(c2)[1], *sc)
    return outs  # gw, gwold, gphi, gpsi, gc


def acoustic_forward(NX, NY, NSTEP, dt, hx, hy, sigma, tau, c, srci, srcj, srcv, rcvi, rcvj, mpi_convention=False,
                     which="oracle", kernel=1):
    """-> (u_hist[(NSTEP+1), NX+2, NY+2], rcvv[(NSTEP+1), nrcv]); kernel=0 (PropagatorKernel=0, oracle only) returns
    (u_hist, upre_hist, rcvv) -- upre_hist holds every step's pre-injection output."""
    N = (NX + 2) * (NY + 2)
    if kernel == 0:
        assert which == "oracle"
        srci, psi_ = _i(srci); srcj, psj_ = _i(srcj); rcvi, pri_ = _i(rcvi); rcvj, prj_ = _i(rcvj)
        srcv = np.ascontiguousarray(srcv, dtype=np.float64)
        assert srcv.shape[0] >= NSTEP and srcv.shape[1] == len(srci)
        u, up = np.empty((NSTEP + 1) * N), np.empty((NSTEP + 1) * N)
        rcvv = np.empty((NSTEP + 1, len(rcvi)))
        lib().orc_acoustic_forward_k0(_i64(NX), _i64(NY), _i64(NSTEP), C.c_double(dt), C.c_double(hx), C.c_double(hy),
                                      _d(sigma)[1], _d(tau)[1], _d(c)[1], C.c_int(int(mpi_convention)),
                                      _i64(len(srci)), psi_, psj_, srcv.ctypes.data_as(_dp), _i64(len(rcvi)), pri_,
                                      prj_, u.ctypes.data_as(_dp), up.ctypes.data_as(_dp), rcvv.ctypes.data_as(_dp))
        return u.reshape(NSTEP + 1, NX + 2, NY + 2), up.reshape(NSTEP + 1, NX + 2, NY + 2), rcvv
    srci, psi_ = _i(srci)
    srcj, psj_ = _i(srcj)
    rcvi, pri_ = _i(rcvi)
    rcvj, prj_ = _i(rcvj)
    srcv = np.ascontiguousarray(srcv, dtype=np.float64)
    assert srcv.shape[0] >= NSTEP and srcv.shape[1] == len(srci)
    u = np.empty((NSTEP + 1) * N)
    rcvv = np.empty((NSTEP + 1, len(rcvi)))
    if which == "oracle":
        lib().orc_acoustic_forward(_i64(NX), _i64(NY), _i64(NSTEP), C.c_double(dt), C.c_double(hx), C.c_double(hy),
                                   _d(sigma)[1], _d(tau)[1], _d(c)[1], C.c_int(int(mpi_convention)),
                                   _i64(len(srci)), psi_, psj_, srcv.ctypes.data_as(_dp), _i64(len(rcvi)), pri_, prj_,
                                   u.ctypes.data_as(_dp), rcvv.ctypes.data_as(_dp))
    else:
        assert not mpi_convention
        ref_lib().ref_drv_acoustic_forward(_i64(NX), _i64(NY), _i64(NSTEP), C.c_double(dt), C.c_double(hx),
                                           C.c_double(hy), _d(sigma)[1], _d(tau)[1], _d(c)[1], _i64(len(srci)), psi_,
                                           psj_, srcv.ctypes.data_as(_dp), _i64(len(rcvi)), pri_, prj_,
                                           u.ctypes.data_as(_dp), rcvv.ctypes.data_as(_dp))
    return u.reshape(NSTEP + 1, NX + 2, NY + 2), rcvv


def acoustic_misfit_grad(NX, NY, NSTEP, dt, hx, hy, sigma, tau, c, srci, srcj, rcvi, rcvj, obs, u_hist,
                         mpi_convention=False, which="oracle", upre_hist=None):
    """-> (loss, grad_c[NX+2, NY+2], grad_srcv[NSTEP, nsrc]); with upre_hist: the PropagatorKernel=0 reverse sweep."""
    N = (NX + 2) * (NY + 2)
    if upre_hist is not None:
        assert which == "oracle"
        srci, psi_ = _i(srci); srcj, psj_ = _i(srcj); rcvi, pri_ = _i(rcvi); rcvj, prj_ = _i(rcvj)
        loss, gc, gs = C.c_double(0.0), np.empty(N), np.empty((NSTEP, len(srci)))
        obs = np.ascontiguousarray(obs, dtype=np.float64)
        u_hist = np.ascontiguousarray(u_hist, dtype=np.float64)
        upre_hist = np.ascontiguousarray(upre_hist, dtype=np.float64)
        lib().orc_acoustic_misfit_grad_k0(_i64(NX), _i64(NY), _i64(NSTEP), C.c_double(dt), C.c_double(hx),
                                          C.c_double(hy), _d(sigma)[1], _d(tau)[1], _d(c)[1],
                                          C.c_int(int(mpi_convention)), _i64(len(srci)), psi_, psj_, _i64(len(rcvi)),
                                          pri_, prj_, obs.ctypes.data_as(_dp), u_hist.ctypes.data_as(_dp),
                                          upre_hist.ctypes.data_as(_dp), C.byref(loss), gc.ctypes.data_as(_dp),
                                          gs.ctypes.data_as(_dp))
        return loss.value, gc.reshape(NX + 2, NY + 2), gs
    srci, psi_ = _i(srci)
    srcj, psj_ = _i(srcj)
    rcvi, pri_ = _i(rcvi)
    rcvj, prj_ = _i(rcvj)
    loss = C.c_double(0.0)
    gc = np.empty(N)
    gs = np.empty((NSTEP, len(srci)))
    obs = np.ascontiguousarray(obs, dtype=np.float64)
    u_hist = np.ascontiguousarray(u_hist, dtype=np.float64)
    if which == "oracle":
        lib().orc_acoustic_misfit_grad(_i64(NX), _i64(NY), _i64(NSTEP), C.c_double(dt), C.c_double(hx),
                                       C.c_double(hy), _d(sigma)[1], _d(tau)[1], _d(c)[1],
                                       C.c_int(int(mpi_convention)), _i64(len(srci)), psi_, psj_, _i64(len(rcvi)),
                                       pri_, prj_, obs.ctypes.data_as(_dp), u_hist.ctypes.data_as(_dp),
                                       C.byref(loss), gc.ctypes.data_as(_dp), gs.ctypes.data_as(_dp))
    else:
        assert not mpi_convention
        ref_lib().ref_drv_acoustic_gradient(_i64(NX), _i64(NY), _i64(NSTEP), C.c_double(dt), C.c_double(hx),
                                            C.c_double(hy), _d(sigma)[1], _d(tau)[1], _d(c)[1], _i64(len(srci)),
                                            psi_, psj_, _i64(len(rcvi)), pri_, prj_, obs.ctypes.data_as(_dp),
                                            u_hist.ctypes.data_as(_dp), C.byref(loss), gc.ctypes.data_as(_dp),
                                            gs.ctypes.data_as(_dp))
    return loss.value, gc.reshape(NX + 2, NY + 2), gs


def ref_mpi_acoustic_forward(NX, NY, n, NSTEP, dt, hx, hy, sigma, tau, c2g, srci, srcj, srcv, nthreads=1):
    """Block-decomposed reference loop (MPI ranks emulated by threads).  -> u[(NSTEP+1), NX, NY]"""
    srci, psi_ = _i(srci)
    srcj, psj_ = _i(srcj)
    srcv = np.ascontiguousarray(srcv, dtype=np.float64)
    u = np.zeros((NSTEP + 1) * NX * NY)
    ref_lib().ref_drv_mpi_acoustic_forward(_i64(NX), _i64(NY), _i64(n), _i64(NSTEP), C.c_double(dt), C.c_double(hx),
                                           C.c_double(hy), _d(sigma)[1], _d(tau)[1], _d(c2g)[1], _i64(len(srci)), psi_,
                                           psj_, srcv.ctypes.data_as(_dp), u.ctypes.data_as(_dp), C.c_int(nthreads))
    return u.reshape(NSTEP + 1, NX, NY)


def ref_mpi_acoustic_gradient(NX, NY, n, NSTEP, dt, hx, hy, sigma, tau, c2g, srci, srcj, rcvi, rcvj, obs, u,
                              nthreads=1):
    """-> (loss, grad_c2[NX, NY], grad_srcv[NSTEP, nsrc]) with the reference's per-block backward body."""
    srci, psi_ = _i(srci)
    srcj, psj_ = _i(srcj)
    rcvi, pri_ = _i(rcvi)
    rcvj, prj_ = _i(rcvj)
    obs = np.ascontiguousarray(obs, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    loss = C.c_double(0.0)
    gc = np.zeros(NX * NY)
    gs = np.zeros((NSTEP, len(srci)))
    ref_lib().ref_drv_mpi_acoustic_gradient(_i64(NX), _i64(NY), _i64(n), _i64(NSTEP), C.c_double(dt), C.c_double(hx),
                                            C.c_double(hy), _d(sigma)[1], _d(tau)[1], _d(c2g)[1], _i64(len(srci)),
                                            psi_, psj_, _i64(len(rcvi)), pri_, prj_, obs.ctypes.data_as(_dp),
                                            u.ctypes.data_as(_dp), C.byref(loss), gc.ctypes.data_as(_dp),
                                            gs.ctypes.data_as(_dp), C.c_int(nthreads))
    return loss.value, gc.reshape(NX, NY), gs


def ref_threads():
    return int(ref_lib().ref_max_threads())


# ----------------------------------------------------------------------------------------------
# elastic
# ----------------------------------------------------------------------------------------------
def elastic_dims(variant, NX, NY):
    H, W = _i64(0), _i64(0)
    lib().orc_elastic_dims(C.c_int(variant), _i64(NX), _i64(NY), C.byref(H), C.byref(W))
    return H.value, W.value


def elastic_forward(variant, NX, NY, NSTEP, dt, dx, dy, ax, bx, ay, by, rho, lam, mu, srci, srcj, srctype, srcv,
                    rcvi, rcvj, rcvtype, want_hist=False):
    """-> (rcvv[nrcv, NSTEP+1], hist[5, NSTEP+1, H, W] or None); rho/lam/mu are H x W (padded) arrays."""
    H, W = elastic_dims(variant, NX, NY)
    srci, p1 = _i(srci)
    srcj, p2 = _i(srcj)
    srctype, p3 = _i(srctype)
    rcvi, p4 = _i(rcvi)
    rcvj, p5 = _i(rcvj)
    rcvtype, p6 = _i(rcvtype)
    srcv = np.ascontiguousarray(srcv, dtype=np.float64)
    assert srcv.shape[0] >= NSTEP and srcv.shape[1] == len(srci)
    rcvv = np.zeros((len(rcvi), NSTEP + 1))
    hist = np.zeros((5, NSTEP + 1, H, W)) if want_hist else None
    lib().orc_elastic_forward(C.c_int(variant), _i64(NX), _i64(NY), _i64(NSTEP), C.c_double(dt), C.c_double(dx),
                              C.c_double(dy), _d(ax)[1], _d(bx)[1], _d(ay)[1], _d(by)[1], _d(rho)[1], _d(lam)[1],
                              _d(mu)[1], _i64(len(srci)), p1, p2, p3, srcv.ctypes.data_as(_dp), _i64(len(rcvi)), p4,
                              p5, p6, rcvv.ctypes.data_as(_dp), hist.ctypes.data_as(_dp) if want_hist else None)
    return rcvv, hist


def elastic_misfit_grad(variant, NX, NY, NSTEP, dt, dx, dy, ax, bx, ay, by, rho, lam, mu, srci, srcj, srctype, srcv,
                        rcvi, rcvj, rcvtype, obs):
    """-> dict(loss, rcvv, grad_rho, grad_lam, grad_mu [H,W], grad_srcv [NSTEP,nsrc])"""
    H, W = elastic_dims(variant, NX, NY)
    srci, p1 = _i(srci)
    srcj, p2 = _i(srcj)
    srctype, p3 = _i(srctype)
    rcvi, p4 = _i(rcvi)
    rcvj, p5 = _i(rcvj)
    rcvtype, p6 = _i(rcvtype)
    srcv = np.ascontiguousarray(srcv, dtype=np.float64)
    obs = np.ascontiguousarray(obs, dtype=np.float64)
    loss = C.c_double(0)
    rcvv = np.zeros((len(rcvi), NSTEP + 1))
    gr, gl, gm = np.zeros((H, W)), np.zeros((H, W)), np.zeros((H, W))
    gs = np.zeros((NSTEP, len(srci)))
    lib().orc_elastic_misfit_grad(C.c_int(variant), _i64(NX), _i64(NY), _i64(NSTEP), C.c_double(dt), C.c_double(dx),
                                  C.c_double(dy), _d(ax)[1], _d(bx)[1], _d(ay)[1], _d(by)[1], _d(rho)[1], _d(lam)[1],
                                  _d(mu)[1], _i64(len(srci)), p1, p2, p3, srcv.ctypes.data_as(_dp), _i64(len(rcvi)),
                                  p4, p5, p6, obs.ctypes.data_as(_dp), C.byref(loss), rcvv.ctypes.data_as(_dp),
                                  gr.ctypes.data_as(_dp), gl.ctypes.data_as(_dp), gm.ctypes.data_as(_dp),
                                  gs.ctypes.data_as(_dp))
    return dict(loss=loss.value, rcvv=rcvv, grad_rho=gr, grad_lam=gl, grad_mu=gm, grad_srcv=gs)


def ref_elastic(variant, NX, NY, NSTEP, dt, dx, dy, ax, bx, ay, by, rho, lam, mu, srci, srcj, srctype, srcv, rcvi, rcvj,
                rcvtype, obs=None, want_grad=True, want_hist=False, block=None):
    """The reference's elastic graph recorded and differentiated op by op with ITS OWN gather / scatter_add /
    scatter_nd / add_source / get_receive bodies (oracle/ref_graph.inc).  variant 0 = src/Core.jl (S), 1 =
    src/MPIElastic.jl (M) on `block` = (n1, n2)-cell blocks (default: one block) with emulated halo exchanges.
    -> dict(loss, rcvv[nrcv, NSTEP+1], hist[5, NSTEP+1, H, W] | None, grad_rho, grad_lam, grad_mu [H, W], grad_srcv)"""
    H, W = elastic_dims(variant, NX, NY)
    srci, p1 = _i(srci); srcj, p2 = _i(srcj); srctype, p3 = _i(srctype)
    rcvi, p4 = _i(rcvi); rcvj, p5 = _i(rcvj); rcvtype, p6 = _i(rcvtype)
    srcv = np.ascontiguousarray(np.asarray(srcv, dtype=np.float64)[:NSTEP])
    nsrc, nrcv = len(srci), len(rcvi)
    assert srcv.shape == (NSTEP, nsrc)
    loss = C.c_double(0.0)
    rcvv = np.zeros((nrcv, NSTEP + 1))
    hist = np.zeros((5, NSTEP + 1, H, W)) if want_hist else None
    gr, gl, gm, gs = np.zeros((H, W)), np.zeros((H, W)), np.zeros((H, W)), np.zeros((NSTEP, nsrc))
    obs_p = None
    if obs is not None:
        obs = np.ascontiguousarray(obs, dtype=np.float64)
        assert obs.shape == rcvv.shape
        obs_p = obs.ctypes.data_as(_dp)
    common = [C.c_double(dt), C.c_double(dx), C.c_double(dy), _d(ax)[1], _d(bx)[1], _d(ay)[1], _d(by)[1], _d(rho)[1],
              _d(lam)[1], _d(mu)[1], _i64(nsrc), p1, p2, p3, srcv.ctypes.data_as(_dp), _i64(nrcv), p4, p5, p6, obs_p,
              C.c_int(int(want_grad and obs is not None)), C.byref(loss), rcvv.ctypes.data_as(_dp),
              hist.ctypes.data_as(_dp) if want_hist else None, gr.ctypes.data_as(_dp), gl.ctypes.data_as(_dp),
              gm.ctypes.data_as(_dp), gs.ctypes.data_as(_dp)]
    if variant == 0:
        ref_lib().ref_drv_elastic_S(_i64(NX), _i64(NY), _i64(NSTEP), *common)
    else:
        n1, n2 = block or (NX, NY)
        assert NX % n1 == 0 and NY % n2 == 0
        ref_lib().ref_drv_elastic_M(_i64(NX), _i64(NY), _i64(n1), _i64(n2), _i64(NSTEP), *common)
    return dict(loss=loss.value, rcvv=rcvv, hist=hist, grad_rho=gr, grad_lam=gl, grad_mu=gm, grad_srcv=gs)


def ref_acoustic_graph(kernel, NX, NY, NSTEP, dt, hx, hy, sigma, tau, c, srci, srcj, srcv, rcvi, rcvj, obs=None,
                       want_grad=True, want_hist=False):
    """AcousticPropagatorSolver as a graph over the reference's own gather / scatter_nd / scatter_add op bodies,
    differentiated with their backward bodies (oracle/ref_graph.inc).  kernel 0 = `one_step` (Core.jl:528-549),
    2 = `acoustic_one_step_customop_ref` (:504-525).  -> dict(loss, rcvv, u, grad_c, grad_srcv)"""
    N = (NX + 2) * (NY + 2)
    srci, p1 = _i(srci); srcj, p2 = _i(srcj); rcvi, p3 = _i(rcvi); rcvj, p4 = _i(rcvj)
    nsrc, nrcv = len(srci), len(rcvi)
    srcv = np.ascontiguousarray(np.asarray(srcv, dtype=np.float64)[:NSTEP]).reshape(NSTEP, nsrc)
    loss = C.c_double(0.0)
    rcvv = np.zeros((NSTEP + 1, nrcv))
    u = np.zeros((NSTEP + 1, NX + 2, NY + 2)) if want_hist else None
    gc, gs = np.zeros((NX + 2, NY + 2)), np.zeros((NSTEP, nsrc))
    obs_p = None
    if obs is not None:
        obs = np.ascontiguousarray(obs, dtype=np.float64)
        obs_p = obs.ctypes.data_as(_dp)
    ref_lib().ref_drv_acoustic_graph(C.c_int(kernel), _i64(NX), _i64(NY), _i64(NSTEP), C.c_double(dt), C.c_double(hx),
                                     C.c_double(hy), _d(sigma)[1], _d(tau)[1], _d(c)[1], _i64(nsrc), p1, p2,
                                     srcv.ctypes.data_as(_dp), _i64(nrcv), p3, p4, obs_p,
                                     C.c_int(int(want_grad and obs is not None)), C.byref(loss),
                                     rcvv.ctypes.data_as(_dp), u.ctypes.data_as(_dp) if want_hist else None,
                                     gc.ctypes.data_as(_dp), gs.ctypes.data_as(_dp))
    return dict(loss=loss.value, rcvv=rcvv, u=u, grad_c=gc, grad_srcv=gs)


def ref_mpi_acoustic_graph(kernel, NX, NY, block, NSTEP, dt, hx, hy, sigma, tau, c2g, srci, srcj, srcv, rcvi, rcvj,
                           obs=None, want_grad=True, want_hist=False):
    """MPIAcousticPropagatorSolver (src/MPIAcoustic.jl) on Mb x Nb blocks of block=(n1, n2) cells held in one process
    (mpi_halo_exchange emulated), as a graph over the reference's op bodies.  Global unpadded grid; c2g = c^2.
    -> dict(loss, rcvv, u[(NSTEP+1), NX, NY], grad_c2[NX, NY], grad_srcv)"""
    n1, n2 = block
    assert NX % n1 == 0 and NY % n2 == 0
    srci, p1 = _i(srci); srcj, p2 = _i(srcj); rcvi, p3 = _i(rcvi); rcvj, p4 = _i(rcvj)
    nsrc, nrcv = len(srci), len(rcvi)
    srcv = np.ascontiguousarray(np.asarray(srcv, dtype=np.float64)[:NSTEP]).reshape(NSTEP, nsrc)
    loss = C.c_double(0.0)
    rcvv = np.zeros((NSTEP + 1, nrcv))
    u = np.zeros((NSTEP + 1, NX, NY)) if want_hist else None
    gc, gs = np.zeros((NX, NY)), np.zeros((NSTEP, nsrc))
    obs_p = None
    if obs is not None:
        obs = np.ascontiguousarray(obs, dtype=np.float64)
        obs_p = obs.ctypes.data_as(_dp)
    ref_lib().ref_drv_mpi_acoustic_graph(C.c_int(kernel), _i64(NX), _i64(NY), _i64(n1), _i64(n2), _i64(NSTEP),
                                         C.c_double(dt), C.c_double(hx), C.c_double(hy), _d(sigma)[1], _d(tau)[1],
                                         _d(c2g)[1], _i64(nsrc), p1, p2, srcv.ctypes.data_as(_dp), _i64(nrcv), p3, p4,
                                         obs_p, C.c_int(int(want_grad and obs is not None)), C.byref(loss),
                                         rcvv.ctypes.data_as(_dp), u.ctypes.data_as(_dp) if want_hist else None,
                                         gc.ctypes.data_as(_dp), gs.ctypes.data_as(_dp))
    return dict(loss=loss.value, rcvv=rcvv, u=u, grad_c2=gc, grad_srcv=gs)
