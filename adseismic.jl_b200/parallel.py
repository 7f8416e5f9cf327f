"""Multi-GPU host logic: one process per GPU, torch.distributed for the plumbing (rendezvous, handle exchange,
all-reduce of losses / gradients); the per-step halo traffic itself never goes through torch or NCCL -- slab plans
store their edge rows straight into the neighbour's device memory over NVLink (CUDA IPC, see csrc/acoustic.cu).

Two modes, as in the reference:
  * shot parallelism  (src/Utils.jl:300-332, 574-600): shot k (1-based) runs on device k % n_gpu; per-device loss and
    gradient are summed -- here with one all-reduce.
  * domain decomposition (src/MPIAcoustic.jl): the reference splits into M x N square blocks over MPI ranks; here the
    padded grid is cut into 1-D slabs along i (rows are contiguous in memory), one slab per GPU.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from .acoustic import AcousticPlan
from .elastic import ElasticPlan


# ------------------------------------------------------------------------------------------------------------
# partitioning helpers (pure host logic; exercised on CPU with the gloo backend in tests/test_parallel_cpu.py)
# ------------------------------------------------------------------------------------------------------------
def shot_assignment(nshots, world):
    """jobs of each rank, 0-based shot indices.  Reference rule (src/Utils.jl:326): device i (1-based) takes the
    1-based shots k with k % n_gpu == i-1."""
    return [[k - 1 for k in range(1, nshots + 1) if k % world == r] for r in range(world)]


def slab_partition(NX, world, rank):
    """(row0, row1): padded rows [row0, row1) owned by `rank` (adseis_slab_partition)."""
    s = _lib.SlabC()
    _lib.check(_lib.load().adseis_slab_partition(int(NX), int(world), int(rank), C.byref(s)))
    return int(s.row0), int(s.row1)


def slab_geometry(NX, NY, world, rank, halo=1):
    """Local array geometry of a slab, identical to the C side: dict(row0,row1,goff,Hl,own0,own1,ld)."""
    row0, row1 = slab_partition(NX, world, rank)
    lo = halo if rank > 0 else 0
    hi = halo if rank < world - 1 else 0
    ld = (NY + 2 + 15) // 16 * 16
    return dict(row0=row0, row1=row1, goff=row0 - lo, Hl=(row1 - row0) + lo + hi, own0=lo, own1=lo + (row1 - row0),
                ld=ld)


def owned_points(pi, row0, row1, mpi_convention):
    """Mask of the sources/receivers whose padded row lies in [row0,row1) (MPIAcoustic.jl:71-78, 98-104)."""
    gi = np.asarray(pi, dtype=np.int64) + (0 if mpi_convention else -1)
    return (gi >= row0) & (gi < row1)


def plan_window(NSTEP, slots):
    """History window W (snapshots) for a budget of `slots` snapshots incl. 4-plane checkpoints (mirrors
    plan_segments in csrc/acoustic.cu)."""
    if slots >= NSTEP + 1:
        return NSTEP + 1
    for W in range(int(slots), 5, -1):
        nseg = (NSTEP - 1 + (W - 2) - 1) // (W - 2)
        if W + 4 * (nseg - 1) <= slots:
            return W
    raise MemoryError("not enough device memory for a history window of 6 snapshots")


def _dist():
    import torch.distributed as dist
    return dist


def init_process_group(backend=None):
    """Rendezvous from the torchrun environment (RANK / WORLD_SIZE / MASTER_*); returns (rank, world, local_rank)."""
    import torch
    dist = _dist()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, world, local_rank


def all_reduce_sum_(t):
    """In-place sum over ranks of a torch tensor (CUDA under nccl, CPU under gloo)."""
    dist = _dist()
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def all_reduce_scalar(x, op="sum", device=None):
    import torch
    dist = _dist()
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op={"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN}[op])
    return float(t.item())


# ------------------------------------------------------------------------------------------------------------
# shot parallelism
# ------------------------------------------------------------------------------------------------------------
def compute_loss_and_grads_GPU(param, srcs, rcvs, Rs, c, ctx=None, plan_cache=None):
    """compute_loss_and_grads_GPU (src/Utils.jl:300-332) for the acoustic solver: sum over shots of
    sum((rcvv-Rs)^2) and its gradient w.r.t. the velocity model `c`.  Shots are dealt to the ranks of the current
    process group with the reference's round-robin rule; the per-rank partial sums are all-reduced, so every rank
    returns the full (loss, grad).  Works single-process too.  Returns (loss: float, grad: np.ndarray)."""
    import torch
    dist = _dist()
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    ctx = ctx or _lib.default_context()
    jobs = shot_assignment(len(srcs), world)[rank]
    shape = (param.NX, param.NY) if param.mpi_convention else (param.NX + 2, param.NY + 2)
    on_gpu = torch.cuda.is_available()
    dev = torch.device("cuda", torch.cuda.current_device()) if on_gpu else torch.device("cpu")
    gsum = torch.zeros(shape, dtype=torch.float64, device=dev)
    gtmp = torch.empty(shape, dtype=torch.float64, device=dev)
    loss = 0.0
    for k in jobs:
        src, rcv = srcs[k], rcvs[k]
        key = (k,)
        plan = plan_cache.get(key) if plan_cache is not None else None
        if plan is None:
            plan = AcousticPlan(param, src.srci, src.srcj, rcv.rcvi, rcv.rcvj, ctx=ctx)
            if plan_cache is not None:
                plan_cache[key] = plan
        plan.set_model(c)
        plan.set_srcv(src.srcv)
        plan.set_obs(Rs[k])
        plan.gradient()
        loss += plan.loss()
        plan.grad_c(out=gtmp)      # device -> device on the context's stream, synchronised on return
        gsum += gtmp
        if on_gpu:                 # torch's stream must be done with gtmp before the next shot overwrites it
            torch.cuda.current_stream().synchronize()
        rcv.rcvv = plan.rcvv()
        if plan_cache is None:
            plan.close()
    all_reduce_sum_(gsum)          # NCCL over NVLink: the reference sums per-GPU gradients on the host
    loss = all_reduce_scalar(loss, "sum", device=dev)
    return loss, gsum.cpu().numpy()


# ------------------------------------------------------------------------------------------------------------
# domain decomposition (acoustic)
# ------------------------------------------------------------------------------------------------------------
class DomainDecomposedAcoustic:
    """One slab of the acoustic solver per rank of the current process group.  Inputs are always the GLOBAL arrays
    (every rank passes the same model / srcv / obs; each keeps what it owns), results are reduced on request."""

    def __init__(self, param, srci, srcj, rcvi, rcvj, ctx=None, hist_slots=None):
        import torch
        dist = _dist()
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.param = param
        self.ctx = ctx or _lib.default_context()
        self.geo = slab_geometry(param.NX, param.NY, self.world, self.rank)
        self.dev = torch.device("cuda", torch.cuda.current_device())
        plane_bytes = self.geo["Hl"] * self.geo["ld"] * 8
        nsrc, nrcv = len(np.atleast_1d(srci)), len(np.atleast_1d(rcvi))
        if self.world == 1:
            self.plan = AcousticPlan(param, srci, srcj, rcvi, rcvj, ctx=self.ctx)
            return
        if hist_slots is None:
            free_b, _ = self.ctx.mem_info()
            model_bytes = (param.NX + 2) * (param.NY + 2) * 8
            # adjoint accumulators, checkpoints of the C side, full-size gradient buffers, torch / NCCL workspaces
            reserve = (16 * plane_bytes + 6 * model_bytes + (4 * (param.NSTEP + 1) * nrcv + 2 * param.NSTEP * nsrc) * 8
                       + (6 << 30))
            slots = max(0, free_b - reserve) // plane_bytes
            W = plan_window(param.NSTEP, slots)
            hist_slots = int(all_reduce_scalar(W, "min", device=self.dev))   # every rank must use the same window
        self.hist_slots = hist_slots
        slab = (self.rank, self.world, self.geo["row0"], self.geo["row1"])
        self.plan = AcousticPlan(param, srci, srcj, rcvi, rcvj, ctx=self.ctx, slab=slab,
                                 hist_bytes_budget=hist_slots * plane_bytes)
        handles = [None] * self.world
        dist.all_gather_object(handles, self.plan.ipc_export())
        lo = handles[self.rank - 1] if self.rank > 0 else None
        hi = handles[self.rank + 1] if self.rank < self.world - 1 else None
        self.plan.ipc_connect(lo, hi)
        dist.barrier()

    def set_model(self, c):
        self.plan.set_model(c)

    def set_srcv(self, srcv):
        self.plan.set_srcv(srcv)

    def set_obs(self, obs):
        self.plan.set_obs(obs)

    def forward(self):
        self.plan.forward()

    def gradient(self):
        self.plan.gradient()

    def loss(self):
        return all_reduce_scalar(self.plan.loss(), "sum", device=self.dev)

    def rcvv(self):
        import torch
        t = torch.from_numpy(self.plan.rcvv()).to(self.dev)
        return all_reduce_sum_(t).cpu().numpy()   # every receiver is owned by exactly one slab, the others hold 0

    def grad_c(self, reduce=True):
        import torch
        shape = self.plan.model_shape
        t = torch.empty(shape, dtype=torch.float64, device=self.dev)
        self.plan.grad_c(out=t)                   # own rows filled, the rest zero
        if reduce:
            all_reduce_sum_(t)
        return t

    def grad_srcv(self):
        import torch
        t = torch.from_numpy(self.plan.grad_srcv()).to(self.dev)
        return all_reduce_sum_(t).cpu().numpy()

    def close(self):
        self.plan.close()


# ------------------------------------------------------------------------------------------------------------
# domain decomposition (elastic; replaces src/MPIElastic.jl's M x N blocks + 18 mpi_halo_exchange2 per step)
# ------------------------------------------------------------------------------------------------------------
EL_HALO = 2


def elastic_slab_partition(param, world, rank):
    """(row0, row1): rows of the INTERNAL elastic array ((NX+2) padded rows for variant 0, NX + 2*2 ghost rows for
    variant 1) owned by `rank` (adseis_elastic_slab_partition)."""
    s = _lib.SlabC()
    pc = param.to_c()
    _lib.check(_lib.load().adseis_elastic_slab_partition(C.byref(pc), int(world), int(rank), C.byref(s)))
    return int(s.row0), int(s.row1)


def elastic_owned_points(param, pi, row0, row1):
    """Mask of sources/receivers whose internal row lies in [row0,row1): 1-based padded index (variant 0) or
    1-based global index (variant 1, MPIElastic.jl:85-86) -> 0-based internal row."""
    gi = np.asarray(pi, dtype=np.int64) + (-1 if param.variant == 0 else 1)
    return (gi >= row0) & (gi < row1)


class DomainDecomposedElastic:
    """One slab of the elastic solver per rank of the current process group.  Inputs are the GLOBAL arrays on every
    rank; results are reduced on request (receivers / sources are owned by exactly one slab)."""

    def __init__(self, param, srci, srcj, srctype, rcvi, rcvj, rcvtype, ctx=None, hist_slots=None):
        import torch
        dist = _dist()
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.param = param
        self.ctx = ctx or _lib.default_context()
        self.dev = torch.device("cuda", torch.cuda.current_device())
        if self.world == 1:
            self.plan = ElasticPlan(param, srci, srcj, srctype, rcvi, rcvj, rcvtype, ctx=self.ctx)
            return
        self.row0, self.row1 = elastic_slab_partition(param, self.world, self.rank)
        # every rank must use the same history window: probe the slot size with a 1-slot plan-free estimate
        H, W = param.NX + (2 if param.variant == 0 else 4), param.NY + (2 if param.variant == 0 else 4)
        ld = (W + 15) // 16 * 16
        Hl = (self.row1 - self.row0) + (EL_HALO if self.rank > 0 else 0) + (EL_HALO if self.rank < self.world - 1 else 0)
        slot_bytes_max = int(all_reduce_scalar(6 * Hl * ld * 8, "max", device=self.dev))   # 5 planes + compact memories
        if hist_slots is None:
            free_b, _ = self.ctx.mem_info()
            reserve = 40 * Hl * ld * 8 + 6 * H * W * 8 + (6 << 30)
            slots = max(2, (free_b - reserve) // slot_bytes_max) if free_b > reserve else 2
            slots = min(slots, param.NSTEP + 1)
            hist_slots = int(all_reduce_scalar(slots, "min", device=self.dev))
        self.hist_slots = int(hist_slots)
        slab = (self.rank, self.world, self.row0, self.row1)
        # the C side sizes the window as budget // slot_bytes: ask for exactly hist_slots slots of MY slot size
        probe = ElasticPlan(param, [], [], [], [], [], [], ctx=self.ctx, slab=slab, hist_bytes_budget=1)
        my_slot_bytes = probe.info()["slot_doubles"] * 8
        probe.close()
        self.plan = ElasticPlan(param, srci, srcj, srctype, rcvi, rcvj, rcvtype, ctx=self.ctx, slab=slab,
                                hist_bytes_budget=self.hist_slots * my_slot_bytes)
        assert self.plan.info()["hist_slots"] == self.hist_slots, (self.plan.info(), self.hist_slots)
        handles = [None] * self.world
        dist.all_gather_object(handles, self.plan.ipc_export())
        lo = handles[self.rank - 1] if self.rank > 0 else None
        hi = handles[self.rank + 1] if self.rank < self.world - 1 else None
        self.plan.ipc_connect(lo, hi)
        dist.barrier()

    def set_model(self, rho, lam, mu):
        self.plan.set_model(rho, lam, mu)

    def set_srcv(self, srcv):
        self.plan.set_srcv(srcv)

    def set_obs(self, obs):
        self.plan.set_obs(obs)

    def forward(self):
        self.plan.forward()

    def gradient(self, material_grads=True):
        self.plan.gradient(material_grads)

    def _sum(self, a):
        import torch
        t = torch.from_numpy(np.ascontiguousarray(a)).to(self.dev)
        return all_reduce_sum_(t).cpu().numpy()

    def loss(self):
        return all_reduce_scalar(self.plan.loss(), "sum", device=self.dev)

    def rcvv(self):
        return self._sum(self.plan.rcvv())          # every receiver is owned by exactly one slab, the others hold 0

    def grad_srcv(self):
        return self._sum(self.plan.grad_srcv())

    def grads(self, reduce=True):
        """(grad_rho, grad_lambda, grad_mu): own rows filled, the rest zero; summed over ranks when reduce."""
        out = (self.plan.grad_rho(), self.plan.grad_lambda(), self.plan.grad_mu())
        return tuple(self._sum(a) for a in out) if (reduce and self.world > 1) else out

    def close(self):
        self.plan.close()


def bench_domain_decomposed(A, w, args, rank, world, local_rank):
    """bench.py at N > 1: the C4 workload slab-partitioned over the ranks (strong scaling).  Returns the JSON dict
    (meaningful on rank 0)."""
    import json
    import time
    import torch
    import bench as B
    init_process_group("nccl")
    dist = _dist()
    ctx = A.Context(local_rank)
    p = A.AcousticPropagatorParams(PropagatorKernel=1, NX=w["NX"], NY=w["NY"], NSTEP=w["NSTEP"], DELTAX=w["DELTAX"], DELTAY=w["DELTAY"],
                                   DELTAT=w["DELTAT"], Rcoef=w["Rcoef"], vp_ref=w["vp_ref"],
                                   NPOINTS_PML=w["NPOINTS_PML"], mpi_convention=True)
    srcv_np = (A.Ricker(p, 100.0, 500.0) * 1e6).reshape(-1, 1)
    dd = DomainDecomposedAcoustic(p, w["srci"], w["srcj"], w["rcvi"], w["rcvj"], ctx=ctx)
    nrcv = len(w["rcvi"])
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_c2, h_srcv = pin(w["c2"]), pin(srcv_np)
    dd.set_model(w["c2_background"]); dd.set_srcv(srcv_np); dd.forward()
    h_obs = pin(dd.rcvv())
    dd.set_model(h_c2.numpy()); dd.set_srcv(h_srcv.numpy()); dd.set_obs(h_obs.numpy())
    for _ in range(args.warmup):
        dd.gradient()
    ctx.sync(); dist.barrier()
    clocks = B.ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    l0 = ctx.launch_count()
    ctx.timer_start()
    for _ in range(args.steps):
        dd.gradient()
    ms = ctx.timer_stop_ms()
    dist.barrier()
    ms = all_reduce_scalar(ms, "max", device=dd.dev)
    launches = ctx.launch_count() - l0
    tm, info = dd.plan.timings(), dd.plan.info()
    loss = dd.loss()
    sec = ms / 1e3 / args.steps
    cells = w["NX"] * w["NY"] * (w["NSTEP"] - 1)
    # e2e: host model / srcv / obs in, loss + (sharded) gradient rows out, every step
    rows = dd.geo["row1"] - dd.geo["row0"]
    h_grad = torch.empty((w["NX"], w["NY"]), dtype=torch.float64).pin_memory()
    ctx.sync(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        dd.set_model(h_c2.numpy()); dd.set_srcv(h_srcv.numpy()); dd.set_obs(h_obs.numpy())
        dd.gradient()
        loss_e2e = dd.loss()
        dd.plan.grad_c(out=h_grad.numpy())
    ctx.sync(); dist.barrier()
    sec_e2e = all_reduce_scalar(time.perf_counter() - t0, "max", device=dd.dev) / args.steps
    clk = clocks.stop() if rank == 0 else None
    peak, peak_src = B.measured_peaks()
    # per-GPU algorithmic bytes of the dominant kernel: this rank's rows
    frac_rows = rows / float(w["NX"] + 2)
    ab = B.algorithmic_bytes(w)
    adj_us = tm["adjoint_ms"] * 1e3 / max(tm["adjoint_launches"], 1)
    achieved = ab["adjoint"] * frac_rows / (adj_us * 1e-6) / 1e9
    roof = dict(bound="hbm", kernel="ac_adj_kernel (+ halo exchange, per GPU, rank 0)", achieved=achieved, peak=peak,
                unit="GB/s", frac=achieved / peak, traffic=None, peak_source=peak_src,
                bytes_per_launch=ab["adjoint"] * frac_rows, us_per_launch=adj_us,
                note="launch duration includes the per-step peer halo exchange that follows every kernel")
    cfg = dict(workload=w["name"], grid=[w["NX"], w["NY"]], nstep=w["NSTEP"], shots=1, dx=w["DELTAX"], dt=w["DELTAT"],
               npml=w["NPOINTS_PML"], nrcv=nrcv, parallelism="slab domain decomposition x%d (NVLink peer halo rows)" % world,
               history_slots=info["hist_slots"], segments=info["segments"],
               recomputed_forward_steps=info["recomputed_steps"],
               l2_policy="working set far exceeds L2; no explicit flush")
    out = dict(metric=B.METRIC, value=cells / sec / 1e9, unit=B.UNIT, n_gpus=world, steps=args.steps,
               warmup=args.warmup, ms_per_step=sec * 1e3, higher_is_better=True, scaling="strong", vs_baseline=None,
               dtype="f64", data="synthetic", config=cfg, clocks=clk,
               e2e=dict(value=cells / sec_e2e / 1e9, unit=B.UNIT,
                        h2d_bytes_per_step=(h_c2.numel() + h_srcv.numel() + h_obs.numel()) * 8 * world,
                        d2h_bytes_per_step=(h_grad.numel() * 8 + 8) * world, ms_per_step=sec_e2e * 1e3),
               gpu_launches=launches * world, roofline=roof, loss=loss, loss_e2e=loss_e2e)
    dd.close()
    return out
