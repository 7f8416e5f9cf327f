"""Host-side mirror of the reference's public structs (src/Struct.jl) -- same field names, defaults and index
conventions (1-based source/receiver indices into the padded grid; srctype/rcvtype 0..4 = vx, vy, sxx, syy, sxy,
src/Struct.jl:44-48)."""
import math
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _lib


@dataclass
class AcousticPropagatorParams:
    """src/Struct.jl:82-121, same defaults.  `PropagatorKernel` keeps the reference's meaning: 0 = the TF-op scheme
    (phi, psi driven by the NEW wavefield, Core.jl:528-549), 1 = the custom-op scheme (phi, psi from the OLD
    wavefield, AcousticOneStepCpu.h), 2 = its op-free twin (numerically scheme 1).  Slab decomposition needs 1."""
    NX: int = 101
    NY: int = 641
    NSTEP: int = 2000 * 2
    DELTAX: float = 10.0
    DELTAY: float = 10.0
    DELTAT: float = 2.0e-3 / 2
    USE_PML_XMIN: bool = True
    USE_PML_XMAX: bool = True
    USE_PML_YMIN: bool = True
    USE_PML_YMAX: bool = True
    NPOINTS_PML: int = 12
    NPOWER: int = 2
    damping_x: Optional[float] = None
    damping_y: Optional[float] = None
    Rcoef: float = 0.001
    vp_ref: float = 1000.0
    Σx: Optional[np.ndarray] = None
    Σy: Optional[np.ndarray] = None
    IT_DISPLAY: int = 0
    PropagatorKernel: int = 0
    mpi_convention: bool = False  # True: MPIAcousticPropagatorParams inputs (src/MPIAcoustic.jl:3-49)

    def to_c(self):
        return _lib.AcousticParamsC(self.NX, self.NY, self.NSTEP, self.DELTAX, self.DELTAY, self.DELTAT,
                                    int(self.USE_PML_XMIN), int(self.USE_PML_XMAX), int(self.USE_PML_YMIN),
                                    int(self.USE_PML_YMAX), self.NPOINTS_PML, self.Rcoef, self.vp_ref,
                                    int(self.mpi_convention), self.PropagatorKernel)


@dataclass
class ElasticPropagatorParams:
    """src/Struct.jl:4-32 (variant 0) / src/MPIElastic.jl:3-58 on the global grid (variant 1)."""
    NX: int = 101
    NY: int = 641
    NSTEP: int = 2000 * 2
    DELTAX: float = 10.0
    DELTAY: float = 10.0
    DELTAT: float = 2.0e-3 / 2
    f0: float = 5.0
    vp_ref: float = 2000.0
    USE_PML_XMIN: bool = True
    USE_PML_XMAX: bool = True
    USE_PML_YMIN: bool = True
    USE_PML_YMAX: bool = True
    NPOINTS_PML: int = 12
    NPOWER: float = 2.0
    K_MAX_PML: float = 1.0
    ALPHA_MAX_PML: Optional[float] = None  # default 2*pi*(f0/2), src/Struct.jl:25
    Rcoef: float = 0.001
    IT_DISPLAY: int = 0
    variant: int = 0  # 0: ElasticPropagatorSolver (src/Core.jl), 1: MPIElasticPropagatorSolver (src/MPIElastic.jl)

    def __post_init__(self):
        if self.ALPHA_MAX_PML is None:
            self.ALPHA_MAX_PML = 2.0 * math.pi * (self.f0 / 2.0)

    def to_c(self):
        return _lib.ElasticParamsC(self.NX, self.NY, self.NSTEP, self.DELTAX, self.DELTAY, self.DELTAT, self.f0,
                                   self.vp_ref, int(self.USE_PML_XMIN), int(self.USE_PML_XMAX),
                                   int(self.USE_PML_YMIN), int(self.USE_PML_YMAX), self.NPOINTS_PML, self.NPOWER,
                                   self.K_MAX_PML, self.ALPHA_MAX_PML, self.Rcoef, self.variant, 0)

    def model_shape(self):
        return (self.NX + 2, self.NY + 2) if self.variant == 0 else (self.NX, self.NY)


@dataclass
class AcousticSource:
    """src/Struct.jl:123-127: srci, srcj (1-based), srcv [>=NSTEP, nsrc]."""
    srci: np.ndarray
    srcj: np.ndarray
    srcv: np.ndarray

    def __post_init__(self):
        self.srci = _lib.as_i64(self.srci)
        self.srcj = _lib.as_i64(self.srcj)
        self.srcv = _lib.as_f64(self.srcv)
        if self.srcv.ndim == 1:
            self.srcv = self.srcv.reshape(-1, 1)
        assert self.srcv.shape[1] == len(self.srci) == len(self.srcj)


@dataclass
class AcousticReceiver:
    """src/Struct.jl:129-133: rcvi, rcvj (1-based), rcvv [(NSTEP+1), nrcv] filled by SimulatedObservation_."""
    rcvi: np.ndarray
    rcvj: np.ndarray
    rcvv: Optional[np.ndarray] = None

    def __post_init__(self):
        self.rcvi = _lib.as_i64(self.rcvi)
        self.rcvj = _lib.as_i64(self.rcvj)


@dataclass
class ElasticSource:
    """src/Struct.jl:50-55"""
    srci: np.ndarray
    srcj: np.ndarray
    srctype: np.ndarray
    srcv: np.ndarray

    def __post_init__(self):
        self.srci = _lib.as_i64(self.srci)
        self.srcj = _lib.as_i64(self.srcj)
        self.srctype = _lib.as_i64(self.srctype)
        self.srcv = _lib.as_f64(self.srcv)
        if self.srcv.ndim == 1:
            self.srcv = self.srcv.reshape(-1, 1)
        assert self.srcv.shape[1] == len(self.srci) == len(self.srcj) == len(self.srctype)


@dataclass
class ElasticReceiver:
    """src/Struct.jl:57-62: rcvv [nrcv, (NSTEP+1)]"""
    rcvi: np.ndarray
    rcvj: np.ndarray
    rcvtype: np.ndarray
    rcvv: Optional[np.ndarray] = None

    def __post_init__(self):
        self.rcvi = _lib.as_i64(self.rcvi)
        self.rcvj = _lib.as_i64(self.rcvj)
        self.rcvtype = _lib.as_i64(self.rcvtype)
