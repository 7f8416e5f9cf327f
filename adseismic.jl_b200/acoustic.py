"""Acoustic host API: thin ctypes wrappers over the C ABI with the reference's entry-point names
(`AcousticPropagatorSolver`, `SimulatedObservation_` for Julia's `SimulatedObservation!`, `compute_PML_Params_`).
All numerics run in libadseis_b200.so on the GPU."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, pd, pi, ptr
from .structs import AcousticPropagatorParams, AcousticReceiver, AcousticSource


def compute_PML_Params_(param: AcousticPropagatorParams):
    """compute_PML_Params!(param) (src/Core.jl:622-655): fills param.Σx, param.Σy ((NX+2) x (NY+2))."""
    lib = _lib.load()
    sx, ty = np.empty(param.NX + 2), np.empty(param.NY + 2)
    pc = param.to_c()
    check(lib.adseis_acoustic_pml_profiles(C.byref(pc), pd(sx), pd(ty)))
    Lx, Ly = param.NPOINTS_PML * param.DELTAX, param.NPOINTS_PML * param.DELTAY
    param.damping_x = param.vp_ref / Lx * np.log(1 / param.Rcoef)
    param.damping_y = param.vp_ref / Ly * np.log(1 / param.Rcoef)
    param.Σx = np.repeat(sx[:, None], param.NY + 2, axis=1)
    param.Σy = np.repeat(ty[None, :], param.NX + 2, axis=0)
    return param


class AcousticPlan:
    """Device-resident state for one (params, sources, receivers) triple; reusable across model updates
    (the FWI loop) -- wraps adseis_acoustic_plan_*."""

    def __init__(self, param, srci, srcj, rcvi, rcvj, ctx=None, slab=None, hist_bytes_budget=0):
        self.lib = _lib.load()
        self.ctx = ctx or _lib.default_context()
        self.param = param
        self.srci, self.srcj = _lib.as_i64(srci), _lib.as_i64(srcj)
        self.rcvi, self.rcvj = _lib.as_i64(rcvi), _lib.as_i64(rcvj)
        self.nsrc, self.nrcv = len(self.srci), len(self.rcvi)
        self.model_shape = (param.NX, param.NY) if param.mpi_convention else (param.NX + 2, param.NY + 2)
        pc = param.to_c()
        h = _lib.vp()
        slab_p = None
        if slab is not None:
            self._slab = _lib.SlabC(*slab)
            slab_p = C.byref(self._slab)
        check(self.lib.adseis_acoustic_plan_create(self.ctx.handle, C.byref(pc), slab_p, self.nsrc, pi(self.srci),
                                                   pi(self.srcj), self.nrcv, pi(self.rcvi), pi(self.rcvj),
                                                   int(hist_bytes_budget), C.byref(h)))
        self.handle = h
        _lib.track(self)

    def set_points(self, srci, srcj, rcvi, rcvj):
        """Replace sources and receivers (next shot, same grid): the device state -- history window, checkpoints,
        model -- is kept.  set_srcv / set_obs must follow."""
        self.srci, self.srcj = _lib.as_i64(srci), _lib.as_i64(srcj)
        self.rcvi, self.rcvj = _lib.as_i64(rcvi), _lib.as_i64(rcvj)
        self.nsrc, self.nrcv = len(self.srci), len(self.rcvi)
        check(self.lib.adseis_acoustic_plan_set_points(self.handle, self.nsrc, pi(self.srci), pi(self.srcj), self.nrcv,
                                                       pi(self.rcvi), pi(self.rcvj)))

    # -- inputs (numpy arrays, torch CUDA tensors or raw device addresses) --
    def set_model(self, c):
        on_dev = int(not isinstance(c, np.ndarray))
        if isinstance(c, np.ndarray):
            c = _lib.as_f64(c)
            assert c.size == self.model_shape[0] * self.model_shape[1], "model has the wrong size"
        self._keep_c = c
        check(self.lib.adseis_acoustic_plan_set_model(self.handle, ptr(c), on_dev))

    def set_srcv(self, srcv, rows=None):
        on_dev = int(not isinstance(srcv, np.ndarray))
        if isinstance(srcv, np.ndarray):
            srcv = _lib.as_f64(srcv)
            rows = srcv.shape[0]
            assert srcv.ndim == 2 and srcv.shape[1] == self.nsrc
        self._keep_s = srcv
        check(self.lib.adseis_acoustic_plan_set_srcv(self.handle, ptr(srcv), int(rows), on_dev))

    def set_obs(self, obs):
        on_dev = int(not isinstance(obs, np.ndarray))
        if isinstance(obs, np.ndarray):
            obs = _lib.as_f64(obs)
            assert obs.shape == (self.param.NSTEP + 1, self.nrcv)
        self._keep_o = obs
        check(self.lib.adseis_acoustic_plan_set_obs(self.handle, ptr(obs), on_dev))

    # -- compute (asynchronous on the ctx stream) --
    def forward(self):
        check(self.lib.adseis_acoustic_plan_forward(self.handle))

    def gradient(self):
        check(self.lib.adseis_acoustic_plan_gradient(self.handle))

    # -- results --
    def _get(self, what, shape, out=None):
        if out is None:
            out = np.empty(shape)
            check(self.lib.adseis_acoustic_plan_get(self.handle, what, ptr(out), 0))
            return out
        check(self.lib.adseis_acoustic_plan_get(self.handle, what, ptr(out), int(not isinstance(out, np.ndarray))))
        return out

    def rcvv(self, out=None):
        return self._get(_lib.GET_RCVV, (self.param.NSTEP + 1, self.nrcv), out)

    def loss(self):
        return float(self._get(_lib.GET_LOSS, (1,))[0])

    def grad_c(self, out=None):
        return self._get(_lib.GET_GRAD_C, self.model_shape, out)

    def owned_model_rows(self):
        """(first, count): rows of the caller's model array this plan's slab owns (the whole model on one GPU)."""
        nrows = self.model_shape[0]
        sl = getattr(self, "_slab", None)
        if sl is None:
            return 0, nrows
        if self.param.mpi_convention:
            r0, r1 = max(int(sl.row0), 1), min(int(sl.row1), self.param.NX + 1)
            return r0 - 1, max(r1 - r0, 0)
        return int(sl.row0), int(sl.row1) - int(sl.row0)

    def grad_c_owned(self, out=None):
        """Only the gradient rows this slab owns (contiguous), for sharded optimisers: (first_row, rows x cols)."""
        first, cnt = self.owned_model_rows()
        return first, self._get(_lib.GET_GRAD_C_OWNED, (cnt, self.model_shape[1]), out)

    def grad_srcv(self, out=None):
        return self._get(_lib.GET_GRAD_SRCV, (self.param.NSTEP, self.nsrc), out)

    def snapshot(self, slot):
        out = np.empty(self.model_shape)
        check(self.lib.adseis_acoustic_plan_get_snapshot(self.handle, int(slot), ptr(out), 0))
        return out

    def info(self):
        a = np.zeros(8, dtype=np.int64)
        check(self.lib.adseis_acoustic_plan_info(self.handle, pi(a)))
        return dict(hist_slots=int(a[0]), segments=int(a[1]), launches=int(a[2]), local_rows=int(a[3]),
                    pitch=int(a[4]), recomputed_steps=int(a[5]), planned_segments=int(a[6]), fast_rows=int(a[7]))

    def timings(self):
        """Device ms / launches of the last forward()/gradient() per phase (CUDA events on the ctx stream)."""
        a = np.zeros(6)
        check(self.lib.adseis_acoustic_plan_timings(self.handle, pd(a)))
        return dict(forward_ms=a[0], recompute_ms=a[1], adjoint_ms=a[2], forward_launches=int(a[3]),
                    recompute_launches=int(a[4]), adjoint_launches=int(a[5]))

    def ipc_export(self):
        """64-byte CUDA IPC handle of this slab plan's device arena (to be all-gathered by the host framework)."""
        buf = C.create_string_buffer(_lib.IPC_HANDLE_BYTES)
        check(self.lib.adseis_acoustic_plan_ipc_export(self.handle, C.cast(buf, C.c_void_p)))
        return bytes(buf.raw)

    def ipc_connect(self, handle_lo, handle_hi):
        lo = C.create_string_buffer(handle_lo, _lib.IPC_HANDLE_BYTES) if handle_lo is not None else None
        hi = C.create_string_buffer(handle_hi, _lib.IPC_HANDLE_BYTES) if handle_hi is not None else None
        check(self.lib.adseis_acoustic_plan_ipc_connect(self.handle, C.cast(lo, C.c_void_p) if lo else None,
                                                        C.cast(hi, C.c_void_p) if hi else None))

    def close(self):
        if getattr(self, "handle", None) is not None:
            self.lib.adseis_acoustic_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class AcousticPropagator:
    """Result of AcousticPropagatorSolver (src/Struct.jl:140-147).  The reference returns a lazy TF graph; here the
    solver call records its inputs and the wavefield history `u` is computed on first access."""

    def __init__(self, param, src, c, ctx=None):
        self.param, self.src, self.c, self.ctx = param, src, c, ctx
        self._u = None

    @property
    def u(self):
        if self._u is None:
            self._u = acoustic_forward(self.param, self.src, self.c, None, want_history=True, ctx=self.ctx)[1]
        return self._u


def AcousticPropagatorSolver(param: AcousticPropagatorParams, src: AcousticSource, c, ctx=None):
    """src/Core.jl:562-620.  `c` is the velocity on the padded grid ((NX+2) x (NY+2)); it is squared inside."""
    compute_PML_Params_(param)
    return AcousticPropagator(param, src, np.asarray(c, dtype=np.float64), ctx)


def acoustic_forward(param, src, c, rcv, want_history=False, ctx=None):
    """One-call forward through adseis_acoustic_forward (host buffers in, host buffers out).
    -> (rcvv [(NSTEP+1), nrcv] or None, u [(NSTEP+1), rows, cols] or None)"""
    lib = _lib.load()
    ctx = ctx or _lib.default_context()
    c = _lib.as_f64(c)
    pc = param.to_c()
    nrcv = 0 if rcv is None else len(rcv.rcvi)
    rcvv = np.empty((param.NSTEP + 1, nrcv)) if nrcv else None
    shape = (param.NX, param.NY) if param.mpi_convention else (param.NX + 2, param.NY + 2)
    assert c.size == shape[0] * shape[1]
    u = np.empty((param.NSTEP + 1,) + shape) if want_history else None
    check(lib.adseis_acoustic_forward(ctx.handle, C.byref(pc), pd(c), len(src.srci), pi(src.srci), pi(src.srcj),
                                      pd(src.srcv), src.srcv.shape[0], nrcv, pi(rcv.rcvi) if nrcv else None,
                                      pi(rcv.rcvj) if nrcv else None, pd(rcvv), pd(u)))
    return rcvv, u


def SimulatedObservation_(ap: AcousticPropagator, rcv: AcousticReceiver):
    """SimulatedObservation!(ap, rcv) (src/Core.jl:726-730): rcv.rcvv[s, r] = u[s, rcvi_r, rcvj_r]."""
    rcv.rcvv, _ = acoustic_forward(ap.param, ap.src, ap.c, rcv, want_history=False, ctx=ap.ctx)
    return rcv.rcvv


def acoustic_misfit_grad(param, src, c, rcv, obs, ctx=None):
    """loss = sum((rcvv-obs)^2) (src/Utils.jl:308) and its gradients w.r.t. c and srcv via
    adseis_acoustic_misfit_grad.  -> dict(loss, rcvv, grad_c, grad_srcv)"""
    lib = _lib.load()
    ctx = ctx or _lib.default_context()
    c = _lib.as_f64(c)
    obs = _lib.as_f64(obs)
    pc = param.to_c()
    nrcv, nsrc = len(rcv.rcvi), len(src.srci)
    assert obs.shape == (param.NSTEP + 1, nrcv)
    loss = C.c_double(0.0)
    rcvv = np.empty((param.NSTEP + 1, nrcv))
    gc = np.empty(c.shape)
    gs = np.empty((param.NSTEP, nsrc))
    check(lib.adseis_acoustic_misfit_grad(ctx.handle, C.byref(pc), pd(c), nsrc, pi(src.srci), pi(src.srcj),
                                          pd(src.srcv), src.srcv.shape[0], nrcv, pi(rcv.rcvi), pi(rcv.rcvj), pd(obs),
                                          C.byref(loss), pd(rcvv), pd(gc), pd(gs)))
    return dict(loss=loss.value, rcvv=rcvv, grad_c=gc, grad_srcv=gs)


# ---- op-level wrappers (device pointers; torch CUDA tensors or raw addresses) ------------------------------
def acoustic_one_step(ctx, w, wold, phi, psi, sigma, tau, c2, dt, hx, hy, NX, NY, u, phiout, psiout, stream=None):
    """acoustic_one_step_customop (src/Core.jl:475-487): AcousticOneStepForward on device arrays."""
    check(_lib.load().adseis_op_acoustic_step_fwd(ctx.handle, ptr(w), ptr(wold), ptr(phi), ptr(psi), ptr(sigma),
                                                  ptr(tau), ptr(c2), dt, hx, hy, NX, NY, ptr(u), ptr(phiout),
                                                  ptr(psiout), stream))


def acoustic_one_step_grad(ctx, gw, gwold, gphi, gpsi, gc, gu, gphiout, gpsiout, w, sigma, tau, c2, dt, hx, hy, NX,
                           NY, stream=None):
    check(_lib.load().adseis_op_acoustic_step_bwd(ctx.handle, ptr(gw), ptr(gwold), ptr(gphi), ptr(gpsi), ptr(gc),
                                                  ptr(gu), ptr(gphiout), ptr(gpsiout), ptr(w), ptr(sigma), ptr(tau),
                                                  ptr(c2), dt, hx, hy, NX, NY, stream))
