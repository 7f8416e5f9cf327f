"""ctypes binding of libadseis_b200.so (the C ABI declared in include/adseis.h).

The library is the product: there is no Python/NumPy/torch fallback for any compute entry point.  If the shared
object is missing it is built in-tree with nvcc (sm_100a); if that fails, or no CUDA device is present at call
time, the error is raised to the caller.
"""
import ctypes as C
import os

import numpy as np

from . import _build

i64 = C.c_int64
f64 = C.c_double
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int64)
vp = C.c_void_p

OK, EINVAL, ECUDA, ENOMEM, ESTATE, ECOMM = 0, -1, -2, -3, -4, -5
GET_RCVV, GET_LOSS, GET_GRAD_C, GET_GRAD_SRCV, GET_GRAD_RHO, GET_GRAD_LAMBDA, GET_GRAD_MU = 1, 2, 3, 4, 5, 6, 7
GET_GRAD_C_OWNED = 8
IPC_HANDLE_BYTES = 64


class AdseisError(RuntimeError):
    def __init__(self, code, text):
        super().__init__("adseis error %d: %s" % (code, text))
        self.code = code


class AcousticParamsC(C.Structure):
    _fields_ = [("NX", i64), ("NY", i64), ("NSTEP", i64), ("DELTAX", f64), ("DELTAY", f64), ("DELTAT", f64),
                ("USE_PML_XMIN", C.c_int32), ("USE_PML_XMAX", C.c_int32), ("USE_PML_YMIN", C.c_int32),
                ("USE_PML_YMAX", C.c_int32), ("NPOINTS_PML", i64), ("Rcoef", f64), ("vp_ref", f64),
                ("mpi_convention", C.c_int32), ("PropagatorKernel", C.c_int32)]


class ElasticParamsC(C.Structure):
    _fields_ = [("NX", i64), ("NY", i64), ("NSTEP", i64), ("DELTAX", f64), ("DELTAY", f64), ("DELTAT", f64),
                ("f0", f64), ("vp_ref", f64), ("USE_PML_XMIN", C.c_int32), ("USE_PML_XMAX", C.c_int32),
                ("USE_PML_YMIN", C.c_int32), ("USE_PML_YMAX", C.c_int32), ("NPOINTS_PML", i64), ("NPOWER", f64),
                ("K_MAX_PML", f64), ("ALPHA_MAX_PML", f64), ("Rcoef", f64), ("variant", C.c_int32),
                ("reserved", C.c_int32)]


class SlabC(C.Structure):
    _fields_ = [("rank", C.c_int32), ("nranks", C.c_int32), ("row0", i64), ("row1", i64)]


# every symbol include/adseis.h declares: name -> (restype, argtypes)
_PA, _PE, _PS = C.POINTER(AcousticParamsC), C.POINTER(ElasticParamsC), C.POINTER(SlabC)
SIGNATURES = {
    "adseis_version": (C.c_int, []),
    "adseis_last_error": (C.c_char_p, []),
    "adseis_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "adseis_ctx_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
    "adseis_ctx_destroy": (C.c_int, [vp]),
    "adseis_ctx_sync": (C.c_int, [vp]),
    "adseis_ctx_stream": (C.c_int, [vp, C.POINTER(vp)]),
    "adseis_ctx_launch_count": (C.c_int, [vp, C.POINTER(i64)]),
    "adseis_ctx_timer_start": (C.c_int, [vp]),
    "adseis_ctx_timer_stop_ms": (C.c_int, [vp, dp]),
    "adseis_ctx_mem_info": (C.c_int, [vp, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "adseis_acoustic_pml_profiles": (C.c_int, [_PA, dp, dp]),
    "adseis_slab_partition": (C.c_int, [i64, C.c_int32, C.c_int32, _PS]),
    "adseis_elastic_slab_partition": (C.c_int, [_PE, C.c_int32, C.c_int32, _PS]),
    "adseis_acoustic_plan_create": (C.c_int, [vp, _PA, _PS, i64, ip, ip, i64, ip, ip, C.c_size_t, C.POINTER(vp)]),
    "adseis_acoustic_plan_destroy": (C.c_int, [vp]),
    "adseis_acoustic_plan_set_points": (C.c_int, [vp, i64, ip, ip, i64, ip, ip]),
    "adseis_acoustic_plan_set_model": (C.c_int, [vp, vp, C.c_int]),
    "adseis_acoustic_plan_set_srcv": (C.c_int, [vp, vp, i64, C.c_int]),
    "adseis_acoustic_plan_set_obs": (C.c_int, [vp, vp, C.c_int]),
    "adseis_acoustic_plan_forward": (C.c_int, [vp]),
    "adseis_acoustic_plan_gradient": (C.c_int, [vp]),
    "adseis_acoustic_plan_get": (C.c_int, [vp, C.c_int, vp, C.c_int]),
    "adseis_acoustic_plan_get_snapshot": (C.c_int, [vp, i64, vp, C.c_int]),
    "adseis_acoustic_plan_info": (C.c_int, [vp, ip]),
    "adseis_acoustic_plan_timings": (C.c_int, [vp, dp]),
    "adseis_acoustic_plan_ipc_export": (C.c_int, [vp, vp]),
    "adseis_acoustic_plan_ipc_connect": (C.c_int, [vp, vp, vp]),
    "adseis_acoustic_forward": (C.c_int, [vp, _PA, dp, i64, ip, ip, dp, i64, i64, ip, ip, dp, dp]),
    "adseis_acoustic_misfit_grad": (C.c_int, [vp, _PA, dp, i64, ip, ip, dp, i64, i64, ip, ip, dp, dp, dp, dp, dp]),
    "adseis_op_acoustic_step_fwd": (C.c_int, [vp] + [vp] * 7 + [f64, f64, f64, i64, i64] + [vp] * 3 + [vp]),
    "adseis_op_acoustic_step_bwd": (C.c_int, [vp] + [vp] * 5 + [vp] * 3 + [vp] * 4 + [f64, f64, f64, i64, i64, vp]),
    "adseis_elastic_cpml_profiles": (C.c_int, [_PE, C.c_int, dp, dp]),
    "adseis_elastic_plan_create": (C.c_int, [vp, _PE, _PS, i64, ip, ip, ip, i64, ip, ip, ip, C.c_size_t,
                                             C.POINTER(vp)]),
    "adseis_elastic_plan_destroy": (C.c_int, [vp]),
    "adseis_elastic_plan_set_model": (C.c_int, [vp, vp, vp, vp, C.c_int]),
    "adseis_elastic_plan_set_srcv": (C.c_int, [vp, vp, i64, C.c_int]),
    "adseis_elastic_plan_set_obs": (C.c_int, [vp, vp, C.c_int]),
    "adseis_elastic_plan_forward": (C.c_int, [vp]),
    "adseis_elastic_plan_gradient": (C.c_int, [vp, C.c_int]),
    "adseis_elastic_plan_get": (C.c_int, [vp, C.c_int, vp, C.c_int]),
    "adseis_elastic_plan_get_snapshot": (C.c_int, [vp, C.c_int, i64, vp, C.c_int]),
    "adseis_elastic_plan_info": (C.c_int, [vp, ip]),
    "adseis_elastic_plan_ipc_export": (C.c_int, [vp, vp]),
    "adseis_elastic_plan_ipc_connect": (C.c_int, [vp, vp, vp]),
    "adseis_elastic_forward": (C.c_int, [vp, _PE, dp, dp, dp, i64, ip, ip, ip, dp, i64, i64, ip, ip, ip, dp, dp]),
    "adseis_elastic_misfit_grad": (C.c_int, [vp, _PE, dp, dp, dp, i64, ip, ip, ip, dp, i64, i64, ip, ip, ip, dp, dp,
                                             dp, dp, dp, dp, dp]),
    "adseis_op_add_source_fwd": (C.c_int, [vp] + [vp] * 10 + [vp, vp, vp, vp, i64, i64, i64, vp]),
    "adseis_op_get_receive_fwd": (C.c_int, [vp] + [vp] * 6 + [i64, vp, vp, vp, i64, i64, i64, vp]),
}

_lib = None


def lib_path():
    return _build.LIB


def load():
    """Load (building first if needed) libadseis_b200.so and attach the signatures.  Raises on failure."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.build()  # no-op when up to date; raises if nvcc fails
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != OK:
        raise AdseisError(rc, load().adseis_last_error().decode("utf-8", "replace"))


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def as_i64(a):
    return np.ascontiguousarray(np.asarray(a).reshape(-1), dtype=np.int64)


def ptr(a):
    """void* of a numpy array / torch tensor / raw int address (None -> NULL)."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError("cannot take the address of %r" % type(a))


def pd(a):
    return a.ctypes.data_as(dp) if a is not None else None


def pi(a):
    return a.ctypes.data_as(ip) if a is not None else None


class Context:
    """One GPU, one stream.  `Context()` binds the current device (LOCAL_RANK under torchrun if set)."""

    def __init__(self, device=None):
        lib = load()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0")) if "LOCAL_RANK" in os.environ else -1
        h = vp()
        check(lib.adseis_ctx_create(int(device), C.byref(h)))
        self.handle = h
        self.lib = lib
        track(self)

    def sync(self):
        check(self.lib.adseis_ctx_sync(self.handle))

    def launch_count(self):
        n = i64(0)
        check(self.lib.adseis_ctx_launch_count(self.handle, C.byref(n)))
        return n.value

    def timer_start(self):
        check(self.lib.adseis_ctx_timer_start(self.handle))

    def timer_stop_ms(self):
        ms = f64(0)
        check(self.lib.adseis_ctx_timer_stop_ms(self.handle, C.byref(ms)))
        return ms.value

    def mem_info(self):
        a, b = C.c_size_t(0), C.c_size_t(0)
        check(self.lib.adseis_ctx_mem_info(self.handle, C.byref(a), C.byref(b)))
        return a.value, b.value

    def stream(self):
        s = vp()
        check(self.lib.adseis_ctx_stream(self.handle, C.byref(s)))
        return s.value

    def close(self):
        if self.handle is not None:
            self.lib.adseis_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = None
_live = None      # weak set of plans / contexts, closed in order (plans first) at interpreter exit


def track(obj):
    """Plans and contexts register here so that interpreter shutdown closes plans before their contexts, while the
    CUDA runtime is still alive (finalisers run in arbitrary order otherwise)."""
    global _live
    if _live is None:
        import atexit
        import weakref
        _live = weakref.WeakSet()

        def _close_all():
            objs = list(_live)
            for o in sorted(objs, key=lambda o: isinstance(o, Context)):
                try:
                    o.close()
                except Exception:
                    pass
        atexit.register(_close_all)
    _live.add(obj)


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context()
    return _default_ctx
