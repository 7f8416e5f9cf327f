"""Seeded synthetic inputs of the five BASELINE.json configurations, exactly as SURVEY.md section 8(d) specifies them
(host-side NumPy, setup only).  bench.py times them and the parity tests run them against the CPU oracle, so both see
the same bytes.  Every builder returns a dict with
    kind   : "acoustic" | "elastic"
    param  : AcousticPropagatorParams | ElasticPropagatorParams  (the reference's structs, src/Struct.jl)
    model  : c (acoustic) | (rho, lam, mu) (elastic)  -- in the layout the param's convention expects
    model_obs : the perturbed model the "observed" data are simulated on
    shots  : list of dict(srci, srcj[, srctype], srcv, rcvi, rcvj[, rcvtype])
`nstep` / `shots` arguments shorten a workload for profiling and CPU-sized parity runs without changing its grid."""
import numpy as np

from .structs import AcousticPropagatorParams, ElasticPropagatorParams
from .utils import Ricker, compute_lame_parameters


def c1(nstep=1000, kernel=0):
    """C1 (BASELINE configs[0], examples/demo + docs/src/index.md:84-89): 2-D acoustic single shot on a small layered
    grid, NX=401, NY=133 (padded 403 x 135), 3 layers vp in {1500, 2500, 3500} split along j, one Ricker source at
    (NX/2, 15), 384 receivers i=10..393 at j=15, observed data from vp*(1 + 0.05 N(0,1))."""
    NX, NY = 401, 133
    rng = np.random.default_rng(1234)
    vp = np.empty((NX + 2, NY + 2))
    third = (NY + 2) // 3
    vp[:, :third], vp[:, third:2 * third], vp[:, 2 * third:] = 1500.0, 2500.0, 3500.0
    p = AcousticPropagatorParams(NX=NX, NY=NY, NSTEP=nstep, DELTAX=26.11, DELTAY=26.11, DELTAT=1.59e-3,
                                 vp_ref=float(vp.mean()), PropagatorKernel=kernel)
    srcv = Ricker(p, 30.0, 200.0, 1e6).reshape(-1, 1)
    rcvi = np.arange(10, 394, dtype=np.int64)
    shot = dict(srci=np.array([NX // 2]), srcj=np.array([15]), srcv=srcv, rcvi=rcvi, rcvj=np.full(len(rcvi), 15))
    return dict(name="C1 acoustic %dx%d nt=%d PropagatorKernel=%d" % (NX, NY, nstep, kernel), kind="acoustic", param=p,
                model=vp, model_obs=vp * (1 + 0.05 * rng.standard_normal(vp.shape)), shots=[shot])


def c2(nstep=1000):
    """C2 (configs[1], examples/demo/ElasticWave.jl:8-27): 2-D elastic (variant S) 500 x 500, dx=1, dt=1e-4,
    vp = 3000 (1 + 0.1 layers), vs = vp/1.732, rho = 2800, vp_ref 3300, vx source Ricker(15, 100, 1e6) at the centre,
    200 receivers of type 0/1 at j=20."""
    NX = NY = 500
    rng = np.random.default_rng(1234)
    p = ElasticPropagatorParams(NX=NX, NY=NY, NSTEP=nstep, DELTAX=1.0, DELTAY=1.0, DELTAT=1e-4, vp_ref=3300.0, variant=0)
    layers = np.floor(np.linspace(0, 4, NY + 2, endpoint=False))[None, :] * np.ones((NX + 2, 1))
    vp = 3000.0 * (1 + 0.1 * layers / 4)
    vs, rho = vp / 1.732, np.full_like(vp, 2800.0)
    lam, mu, rho = compute_lame_parameters(vp, vs, rho)
    vpo = vp * (1 + 0.05 * rng.standard_normal(vp.shape))
    lamo, muo, _ = compute_lame_parameters(vpo, vpo / 1.732, rho)
    srcv = Ricker(p, 15.0, 100.0, 1e6).reshape(-1, 1)
    rcvi = np.linspace(50, 450, 200).astype(np.int64)
    shot = dict(srci=np.array([NX // 2]), srcj=np.array([NY // 2]), srctype=np.array([0]), srcv=srcv, rcvi=rcvi,
                rcvj=np.full(200, 20), rcvtype=np.arange(200) % 2)
    return dict(name="C2 elastic(S) %dx%d nt=%d" % (NX, NY, nstep), kind="elastic", param=p, model=(rho, lam, mu),
                model_obs=(rho, lamo, muo), shots=[shot])


def c3(nstep=3000, shots=64):
    """C3 (configs[2]): multi-shot acoustic FWI gradient on a 2000 x 1000 (x by z) Marmousi-shaped model (linear
    gradient + random layered perturbation), `shots` sources evenly on i in [40, 1960] at j=12, 1921 receivers at
    j=12; shots are dealt round-robin k % n_gpu (src/Utils.jl:326) and the gradients all-reduced."""
    NX, NY = 2000, 1000
    rng = np.random.default_rng(1234)
    z = np.arange(NY + 2)[None, :] / (NY + 1.0)
    layers = np.cumsum(rng.standard_normal(40))          # 40 random layers, smooth lateral undulation
    lay = layers[np.minimum((z * 40).astype(int), 39)]
    x = np.arange(NX + 2)[:, None] / (NX + 1.0)
    vp = 1500.0 + 3000.0 * z + 120.0 * lay * (1 + 0.2 * np.sin(6.0 * x)) + 0.0 * x
    vp = np.clip(vp, 1450.0, 4800.0)
    p = AcousticPropagatorParams(NX=NX, NY=NY, NSTEP=nstep, DELTAX=10.0, DELTAY=10.0, DELTAT=1e-3,
                                 vp_ref=float(vp.mean()), PropagatorKernel=1)
    src_i = np.linspace(40, 1960, shots).astype(np.int64)
    rcvi = np.arange(40, 1961, dtype=np.int64)
    srcv = Ricker(p, 30.0, 200.0, 1e6).reshape(-1, 1)
    sh = [dict(srci=np.array([i]), srcj=np.array([12]), srcv=srcv, rcvi=rcvi, rcvj=np.full(len(rcvi), 12)) for i in src_i]
    from scipy.ndimage import gaussian_filter
    return dict(name="C3 acoustic %dx%d nt=%d %d shots" % (NX, NY, nstep, shots), kind="acoustic", param=p,
                model=gaussian_filter(vp, 12.0), model_obs=vp, shots=sh)


def c4(nstep=5000, nx=4096, ny=4096):
    """C4 (configs[3], examples/mpi_acoustic_optimized/MPI_forward.jl:25-38 scaled): acoustic 4096^2, nt=5000, dx=10,
    dt=0.05, c^2 = 1000 with a 2000 square inclusion (centre +- N/8), Rcoef 0.2, source (NX/5, NY/2) Ricker(100, 500),
    receivers j=20..NY-19 at i=NX/5; MPI convention (c given as c^2 on the unpadded grid, unpadded indices)."""
    NX, NY = nx, ny
    p = AcousticPropagatorParams(PropagatorKernel=1, NX=NX, NY=NY, NSTEP=nstep, DELTAX=10.0, DELTAY=10.0, DELTAT=0.05,
                                 Rcoef=0.2, vp_ref=1000.0, NPOINTS_PML=12, mpi_convention=True)
    c2 = np.full((NX, NY), 1000.0)
    cx, cy, wx, wy = NX // 2, NY // 2, NX // 8, NY // 8
    c2[cx - wx - 1:cx + wx, cy - wy - 1:cy + wy] = 2000.0
    rcvj = np.arange(20, NY - 18, dtype=np.int64)
    shot = dict(srci=np.array([NX // 5]), srcj=np.array([NY // 2]), srcv=(Ricker(p, 100.0, 500.0) * 1e6).reshape(-1, 1),
                rcvi=np.full(len(rcvj), NX // 5), rcvj=rcvj)
    return dict(name="C4 acoustic %dx%d nt=%d (mpi_acoustic_optimized analogue)" % (NX, NY, nstep), kind="acoustic",
                param=p, model=c2, model_obs=np.full((NX, NY), 1100.0), shots=[shot])


def c5(nstep=2000, n=2000):
    """C5 (configs[4], examples/mpi_elastic_v2/MPI_forward.jl:11-18): elastic variant M 2000^2, nt=2000, dx=1,
    dt=5e-5, vp=3300, vs=3300/1.732, rho=2800, vx source at (NX/5, NY/2), vx receivers; gradient w.r.t. the source
    time function and lambda."""
    NX = NY = n
    rng = np.random.default_rng(1234)
    p = ElasticPropagatorParams(NX=NX, NY=NY, NSTEP=nstep, DELTAX=1.0, DELTAY=1.0, DELTAT=5e-5, vp_ref=3300.0, variant=1)
    shape = p.model_shape()
    vp = np.full(shape, 3300.0)
    lam, mu, rho = compute_lame_parameters(vp, vp / 1.732, np.full(shape, 2800.0))
    vpo = vp * (1 + 0.03 * rng.standard_normal(shape))
    lamo, muo, _ = compute_lame_parameters(vpo, vpo / 1.732, rho)
    rcvj = np.arange(20, NY - 18, 4, dtype=np.int64)
    shot = dict(srci=np.array([NX // 5]), srcj=np.array([NY // 2]), srctype=np.array([0]),
                srcv=Ricker(p, 50.0, 200.0, 1e6).reshape(-1, 1), rcvi=np.full(len(rcvj), NX // 5 + 40), rcvj=rcvj,
                rcvtype=np.zeros(len(rcvj), dtype=np.int64))
    return dict(name="C5 elastic(M) %dx%d nt=%d (mpi_elastic_v2 analogue)" % (NX, NY, nstep), kind="elastic", param=p,
                model=(rho, lam, mu), model_obs=(rho, lamo, muo), shots=[shot])


BUILDERS = dict(c1=c1, c2=c2, c3=c3, c4=c4, c5=c5)


# ALGORITHMIC bytes per cell-update (fp64, compulsory traffic, perfect stencil reuse)
ACOUSTIC_FWD_B, ACOUSTIC_ADJ_B, ACOUSTIC_PML_EXTRA_B = 32, 56, 32
ELASTIC_FWD_B, ELASTIC_ADJ_B, ELASTIC_PML_EXTRA_B = 104, 192, 24
ELASTIC_ADJ_SRC_ONLY_B = 80          # 5 adjoint fields read + written; no forward state, no accumulators


def algorithmic_bytes(w):
    """Per-launch-set algorithmic bytes of one forward step and one adjoint step over the whole grid (8(d) figures;
    PML-frame cells carry their extra auxiliary-field traffic)."""
    p = w["param"]
    n = p.NPOINTS_PML + 1
    N = p.NX * p.NY
    Np = N - max(p.NX - 2 * n, 0) * max(p.NY - 2 * n, 0)
    if w["kind"] == "acoustic":
        return dict(forward=ACOUSTIC_FWD_B * N + ACOUSTIC_PML_EXTRA_B * Np,
                    adjoint=ACOUSTIC_ADJ_B * N + ACOUSTIC_PML_EXTRA_B * Np)
    return dict(forward=ELASTIC_FWD_B * N + ELASTIC_PML_EXTRA_B * Np, adjoint=ELASTIC_ADJ_B * N + ELASTIC_PML_EXTRA_B * Np,
                adjoint_source_only=ELASTIC_ADJ_SRC_ONLY_B * N + ELASTIC_PML_EXTRA_B * Np)
