"""Elastic host API: ctypes wrappers over the C ABI with the reference's entry-point names
(`ElasticPropagatorSolver`, `SimulatedObservation_`).  All numerics run in libadseis_b200.so on the GPU."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, pd, pi, ptr
from .structs import ElasticPropagatorParams, ElasticReceiver, ElasticSource


def compute_PML_Params(param: ElasticPropagatorParams):
    """compute_PML_Params(param) (src/Core.jl:231-407) -> ax, bx, kx, ay, by, ky with shape (2, NX) / (2, NY)
    (row 0 integer grid, row 1 half grid; K == 1, which the reference ignores anyway, Core.jl:686-693)."""
    lib = _lib.load()
    pc = param.to_c()
    ax, bx = np.empty(2 * param.NX), np.empty(2 * param.NX)
    ay, by = np.empty(2 * param.NY), np.empty(2 * param.NY)
    check(lib.adseis_elastic_cpml_profiles(C.byref(pc), 0, pd(ax), pd(bx)))
    check(lib.adseis_elastic_cpml_profiles(C.byref(pc), 1, pd(ay), pd(by)))
    sh = lambda a, n: a.reshape(2, n)
    return (sh(ax, param.NX), sh(bx, param.NX), np.ones((2, param.NX)), sh(ay, param.NY), sh(by, param.NY),
            np.ones((2, param.NY)))


class ElasticPlan:
    """Device-resident state for one (params, sources, receivers) triple -- wraps adseis_elastic_plan_*."""

    def __init__(self, param, srci, srcj, srctype, rcvi, rcvj, rcvtype, ctx=None, hist_bytes_budget=0, slab=None):
        self.lib = _lib.load()
        self.ctx = ctx or _lib.default_context()
        self.param = param
        self.srci, self.srcj, self.srctype = _lib.as_i64(srci), _lib.as_i64(srcj), _lib.as_i64(srctype)
        self.rcvi, self.rcvj, self.rcvtype = _lib.as_i64(rcvi), _lib.as_i64(rcvj), _lib.as_i64(rcvtype)
        self.nsrc, self.nrcv = len(self.srci), len(self.rcvi)
        self.model_shape = param.model_shape()
        pc = param.to_c()
        h = _lib.vp()
        sl = None
        if slab is not None:   # (rank, nranks, row0, row1): rows of the internal array owned by this GPU
            self._slab = _lib.SlabC(int(slab[0]), int(slab[1]), int(slab[2]), int(slab[3]))
            sl = C.byref(self._slab)
        check(self.lib.adseis_elastic_plan_create(self.ctx.handle, C.byref(pc), sl, self.nsrc, pi(self.srci),
                                                  pi(self.srcj), pi(self.srctype), self.nrcv, pi(self.rcvi),
                                                  pi(self.rcvj), pi(self.rcvtype), int(hist_bytes_budget),
                                                  C.byref(h)))
        self.handle = h
        _lib.track(self)

    def set_model(self, rho, lam, mu):
        arrs = []
        for a in (rho, lam, mu):
            if isinstance(a, np.ndarray):
                a = _lib.as_f64(a)
                assert a.size == self.model_shape[0] * self.model_shape[1], "model has the wrong size"
            arrs.append(a)
        self._keep_m = arrs
        check(self.lib.adseis_elastic_plan_set_model(self.handle, ptr(arrs[0]), ptr(arrs[1]), ptr(arrs[2]),
                                                     int(not isinstance(rho, np.ndarray))))

    def set_srcv(self, srcv, rows=None):
        if isinstance(srcv, np.ndarray):
            srcv = _lib.as_f64(srcv)
            rows = srcv.shape[0]
            assert srcv.ndim == 2 and srcv.shape[1] == self.nsrc
        self._keep_s = srcv
        check(self.lib.adseis_elastic_plan_set_srcv(self.handle, ptr(srcv), int(rows),
                                                    int(not isinstance(srcv, np.ndarray))))

    def set_obs(self, obs):
        if isinstance(obs, np.ndarray):
            obs = _lib.as_f64(obs)
            assert obs.shape == (self.nrcv, self.param.NSTEP + 1)
        self._keep_o = obs
        check(self.lib.adseis_elastic_plan_set_obs(self.handle, ptr(obs), int(not isinstance(obs, np.ndarray))))

    def forward(self):
        check(self.lib.adseis_elastic_plan_forward(self.handle))

    def gradient(self, material_grads=True):
        check(self.lib.adseis_elastic_plan_gradient(self.handle, int(bool(material_grads))))

    def _get(self, what, shape, out=None):
        if out is None:
            out = np.empty(shape)
        check(self.lib.adseis_elastic_plan_get(self.handle, what, ptr(out), int(not isinstance(out, np.ndarray))))
        return out

    def rcvv(self, out=None):
        return self._get(_lib.GET_RCVV, (self.nrcv, self.param.NSTEP + 1), out)

    def loss(self):
        return float(self._get(_lib.GET_LOSS, (1,))[0])

    def grad_srcv(self, out=None):
        return self._get(_lib.GET_GRAD_SRCV, (self.param.NSTEP, self.nsrc), out)

    def grad_rho(self, out=None):
        return self._get(_lib.GET_GRAD_RHO, self.model_shape, out)

    def grad_lambda(self, out=None):
        return self._get(_lib.GET_GRAD_LAMBDA, self.model_shape, out)

    def grad_mu(self, out=None):
        return self._get(_lib.GET_GRAD_MU, self.model_shape, out)

    def snapshot(self, field, slot):
        out = np.empty(self.model_shape)
        check(self.lib.adseis_elastic_plan_get_snapshot(self.handle, int(field), int(slot), ptr(out), 0))
        return out

    def info(self):
        a = np.zeros(8, dtype=np.int64)
        check(self.lib.adseis_elastic_plan_info(self.handle, pi(a)))
        return dict(hist_slots=int(a[0]), segments=int(a[1]), launches=int(a[2]), local_rows=int(a[3]),
                    pitch=int(a[4]), recomputed_steps=int(a[5]), planned_segments=int(a[6]), slot_doubles=int(a[7]))

    def ipc_export(self):
        """64-byte CUDA IPC handle of this slab plan's device arena (to be all-gathered by the host framework)."""
        buf = C.create_string_buffer(_lib.IPC_HANDLE_BYTES)
        check(self.lib.adseis_elastic_plan_ipc_export(self.handle, C.cast(buf, C.c_void_p)))
        return bytes(buf.raw)

    def ipc_connect(self, handle_lo, handle_hi):
        lo = C.create_string_buffer(handle_lo, _lib.IPC_HANDLE_BYTES) if handle_lo is not None else None
        hi = C.create_string_buffer(handle_hi, _lib.IPC_HANDLE_BYTES) if handle_hi is not None else None
        check(self.lib.adseis_elastic_plan_ipc_connect(self.handle, C.cast(lo, C.c_void_p) if lo else None,
                                                       C.cast(hi, C.c_void_p) if hi else None))

    def close(self):
        if getattr(self, "handle", None) is not None:
            self.lib.adseis_elastic_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def elastic_forward(param, src, rho, lam, mu, rcv, want_history=False, ctx=None):
    """adseis_elastic_forward: -> (rcvv [nrcv, NSTEP+1] or None, hist [5, NSTEP+1, rows, cols] or None)."""
    lib = _lib.load()
    ctx = ctx or _lib.default_context()
    rho, lam, mu = _lib.as_f64(rho), _lib.as_f64(lam), _lib.as_f64(mu)
    pc = param.to_c()
    nrcv = 0 if rcv is None else len(rcv.rcvi)
    rcvv = np.empty((nrcv, param.NSTEP + 1)) if nrcv else None
    hist = np.empty((5, param.NSTEP + 1) + param.model_shape()) if want_history else None
    check(lib.adseis_elastic_forward(ctx.handle, C.byref(pc), pd(rho), pd(lam), pd(mu), len(src.srci), pi(src.srci),
                                     pi(src.srcj), pi(src.srctype), pd(src.srcv), src.srcv.shape[0], nrcv,
                                     pi(rcv.rcvi) if nrcv else None, pi(rcv.rcvj) if nrcv else None,
                                     pi(rcv.rcvtype) if nrcv else None, pd(rcvv), pd(hist)))
    return rcvv, hist


class ElasticPropagator:
    """Result of ElasticPropagatorSolver (src/Struct.jl:64-73); field histories are computed on first access."""

    def __init__(self, param, src, rho, lam, mu, ctx=None):
        self.param, self.src, self.rho, self.lam, self.mu, self.ctx = param, src, rho, lam, mu, ctx
        self._hist = None

    def _fields(self):
        if self._hist is None:
            self._hist = elastic_forward(self.param, self.src, self.rho, self.lam, self.mu, None, True, self.ctx)[1]
        return self._hist

    vx = property(lambda self: self._fields()[0])
    vy = property(lambda self: self._fields()[1])
    sigmaxx = property(lambda self: self._fields()[2])
    sigmayy = property(lambda self: self._fields()[3])
    sigmaxy = property(lambda self: self._fields()[4])


def ElasticPropagatorSolver(param: ElasticPropagatorParams, src: ElasticSource, rho, lam, mu, ctx=None):
    """src/Core.jl:31-94 (variant 0) / src/MPIElastic.jl:374-466 on the global grid (variant 1)."""
    return ElasticPropagator(param, src, np.asarray(rho, dtype=np.float64), np.asarray(lam, dtype=np.float64),
                             np.asarray(mu, dtype=np.float64), ctx)


def elastic_misfit_grad(param, src, rho, lam, mu, rcv, obs, material_grads=True, ctx=None):
    """loss = sum((rcvv-obs)^2) and d loss / d (rho, lambda, mu, srcv) via adseis_elastic_misfit_grad."""
    lib = _lib.load()
    ctx = ctx or _lib.default_context()
    rho, lam, mu, obs = _lib.as_f64(rho), _lib.as_f64(lam), _lib.as_f64(mu), _lib.as_f64(obs)
    pc = param.to_c()
    nrcv, nsrc = len(rcv.rcvi), len(src.srci)
    assert obs.shape == (nrcv, param.NSTEP + 1)
    loss = C.c_double(0.0)
    rcvv = np.empty((nrcv, param.NSTEP + 1))
    gs = np.empty((param.NSTEP, nsrc))
    sh = param.model_shape()
    gr, gl, gm = (np.empty(sh), np.empty(sh), np.empty(sh)) if material_grads else (None, None, None)
    check(lib.adseis_elastic_misfit_grad(ctx.handle, C.byref(pc), pd(rho), pd(lam), pd(mu), nsrc, pi(src.srci),
                                         pi(src.srcj), pi(src.srctype), pd(src.srcv), src.srcv.shape[0], nrcv,
                                         pi(rcv.rcvi), pi(rcv.rcvj), pi(rcv.rcvtype), pd(obs), C.byref(loss), pd(rcvv),
                                         pd(gr), pd(gl), pd(gm), pd(gs)))
    return dict(loss=loss.value, rcvv=rcvv, grad_rho=gr, grad_lambda=gl, grad_mu=gm, grad_srcv=gs)
