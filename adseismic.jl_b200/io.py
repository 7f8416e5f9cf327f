"""MAT-file model / shot I/O of the reference (src/IO.jl:4-161), on scipy.io: the shipped Marmousi / BP / layer
fixtures (examples/nn_fwi/models/*.mat, examples/invert_velocity_model/models/*.mat; MAT v5) load into the same
structs.  Conventions as in the reference: `nx`,`ny` in the file are PADDED sizes (NX = nx-2), source / receiver
indices are 1-based into the padded grid, `source[k].vec` has one column per source point and `nt` rows."""
import numpy as np

from .structs import (AcousticPropagatorParams, AcousticReceiver, AcousticSource, ElasticPropagatorParams,
                      ElasticReceiver, ElasticSource)


def _matread(filename):
    import scipy.io as sio
    d = sio.loadmat(filename, squeeze_me=False, struct_as_record=False)
    return {k: v for k, v in d.items() if not k.startswith("__")}


def _scalar(x):
    return np.asarray(x).reshape(-1)[0].item()


def _items(cell):
    """iterate a MATLAB struct array / cell array of structs"""
    for it in np.asarray(cell, dtype=object).reshape(-1):
        while isinstance(it, np.ndarray) and it.dtype == object and it.size == 1:
            it = it.reshape(-1)[0]
        yield it


def _field(it, name):
    return getattr(it, name) if hasattr(it, name) else it[name]


def safe_vec(a):
    """src/IO.jl:163-169"""
    return np.asarray(a, dtype=np.int64).reshape(-1)


def load_params(filename, option="Acoustic", **kwargs):
    """src/IO.jl:4-28 (Acoustic / Elastic; the MPIElastic option maps to ElasticPropagatorParams(variant=1))."""
    d = _matread(filename)
    nx, ny, nt = int(_scalar(d["nx"])), int(_scalar(d["ny"])), int(_scalar(d["nt"]))
    common = dict(NSTEP=nt, DELTAX=float(_scalar(d["dx"])), DELTAY=float(_scalar(d["dy"])), DELTAT=float(_scalar(d["dt"])))
    if option == "Acoustic":
        return AcousticPropagatorParams(NX=nx - 2, NY=ny - 2, **common, **kwargs)
    if option == "Elastic":
        return ElasticPropagatorParams(NX=nx - 2, NY=ny - 2, **common, **kwargs)
    if option == "MPIElastic":
        return ElasticPropagatorParams(NX=nx, NY=ny, variant=1, **common, **kwargs)
    raise ValueError("option should be Acoustic, Elastic or MPIElastic")


def load_acoustic_model(filename, **kwargs):
    """src/IO.jl:101-126 -> (param, vp): vp is the padded (nx, ny) velocity array of the file.  The reference returns
    the closure `src -> AcousticPropagatorSolver(param, src, vp)`; use AcousticPropagatorSolver / fwi.acoustic_misfit
    with these two."""
    d = _matread(filename)
    param = load_params(filename, "Acoustic", **kwargs)
    return param, np.ascontiguousarray(d["vp"], dtype=np.float64)


def load_elastic_model(filename, **kwargs):
    """src/IO.jl:30-48 -> (param, vp, vs, rho) with vp_ref = mean(vp), f0 = file f0 / 2."""
    d = _matread(filename)
    vp = np.ascontiguousarray(d["vp"], dtype=np.float64)
    kw = dict(vp_ref=float(vp.mean()), f0=float(_scalar(d["f0"])) / 2)
    kw.update(kwargs)
    param = load_params(filename, "Elastic", **kw)
    return (param, vp, np.ascontiguousarray(d["vs"], dtype=np.float64),
            np.ascontiguousarray(d["rho"], dtype=np.float64))


def _srcvec(it, nt):
    vec = np.asarray(_field(it, "vec"), dtype=np.float64)
    if vec.ndim == 1 or vec.shape[0] == 1:
        vec = vec.reshape(-1, 1)
    if vec.shape[0] != nt and vec.shape[1] == nt:
        vec = np.ascontiguousarray(vec.T)
    assert vec.shape[0] == nt, "source time function has %d rows, nt = %d" % (vec.shape[0], nt)
    return np.ascontiguousarray(vec)


def load_acoustic_source(filename):
    """src/IO.jl:141-153 -> list of AcousticSource (one per shot)"""
    d = _matread(filename)
    nt = int(_scalar(d["nt"]))
    return [AcousticSource(safe_vec(_field(it, "ix")), safe_vec(_field(it, "iy")), _srcvec(it, nt))
            for it in _items(d["source"])]


def load_acoustic_receiver(filename):
    """src/IO.jl:129-139 -> list of AcousticReceiver"""
    d = _matread(filename)
    return [AcousticReceiver(safe_vec(_field(it, "ix")), safe_vec(_field(it, "iy"))) for it in _items(d["receiver"])]


def load_elastic_source(filename):
    """src/IO.jl:69-92"""
    d = _matread(filename)
    nt = int(_scalar(d["nt"]))
    return [ElasticSource(safe_vec(_field(it, "ix")), safe_vec(_field(it, "iy")), safe_vec(_field(it, "type")),
                          _srcvec(it, nt)) for it in _items(d["source"])]


def load_elastic_receiver(filename):
    """src/IO.jl:50-67"""
    d = _matread(filename)
    return [ElasticReceiver(safe_vec(_field(it, "ix")), safe_vec(_field(it, "iy")), safe_vec(_field(it, "type")))
            for it in _items(d["receiver"])]


def save_model(filename, vp, vs, rho, sources, receivers, dx, dy, dt, nt, f0):
    """Write a model / acquisition file in the reference's MAT schema (what its generate_*.py scripts produce)."""
    import scipy.io as sio
    def pack(seq, keys):
        arr = np.empty((1, len(seq)), dtype=object)
        for k, s in enumerate(seq):
            arr[0, k] = {kk: np.asarray(getattr(s, attr)) for kk, attr in keys}
        return arr
    src_keys = [("ix", "srci"), ("iy", "srcj"), ("vec", "srcv")] + ([("type", "srctype")] if hasattr(sources[0], "srctype") else [])
    rcv_keys = [("ix", "rcvi"), ("iy", "rcvj")] + ([("type", "rcvtype")] if hasattr(receivers[0], "rcvtype") else [])
    sio.savemat(filename, dict(vp=vp, vs=vs, rho=rho, source=pack(sources, src_keys), receiver=pack(receivers, rcv_keys),
                               dx=float(dx), dy=float(dy), dt=float(dt), nx=int(vp.shape[0]), ny=int(vp.shape[1]),
                               nt=int(nt), f0=f0))
